"""The four m x m x n products of an iteration (ry = W rho^T, Qij = ry rinv, H = T rinv^T, grad = G0 + H W;
linearcorex.py:261, :266, :294, :300) on the int8 tcgen05 engine instead of DMMA (`LCX_MM_I8=1`; what large-m
problems such as BASELINE config 4 use).  Same bar as everything else in the FP64-faithful mode: 1e-9 against the
reference's float64 path, clusters bit-exact."""
import ctypes as C

import numpy as np
import pytest

from conftest import load_golden
from test_gpu_parity import RTOL, _check_fit, _fit, assert_close

pytestmark = pytest.mark.gpu


@pytest.fixture
def mm_i8(monkeypatch):
    monkeypatch.setenv("LCX_MM_I8", "1")


@pytest.mark.parametrize("name", ["readme_demo_f64", "big5_l0_f64", "syn_400x300x10_f64", "syn_60x400x8_f64",
                                  "outliers_missing_f64", "big5_l1_f64"])
def test_full_fit_mm_i8(mm_i8, name):
    z, mdl, x = _fit(name, precision="fp64_split")
    _check_fit(z, mdl, x, RTOL)


def _arrays_after_step(monkeypatch, flag, precision, xt_np, w_np, eps):
    """quick moments, direction and one linear trial; returns the arrays the four products feed."""
    import torch
    from linearcorex_b200 import _lib as L
    from linearcorex_b200.corex import _DeviceSession
    if flag is None:
        monkeypatch.delenv("LCX_MM_I8", raising=False)
    else:
        monkeypatch.setenv("LCX_MM_I8", flag)
    sess = _DeviceSession(L.PRECISIONS[precision])
    N, n = xt_np.shape
    m = w_np.shape[0]
    ld = sess.lib.lcx_ld(n)
    xt = torch.zeros((N, ld), dtype=torch.float64, device="cuda")
    xt[:, :n] = torch.from_numpy(xt_np)
    sess.bind(xt, N, n, m, None)
    L.check(sess.lib.lcx_set_w(sess.h, w_np.ctypes.data_as(C.c_void_p), n))
    tc, muj, tang = C.c_double(), C.c_double(), C.c_double()
    L.check(sess.lib.lcx_moments_ns(sess.h, eps, 1, C.byref(tc), C.byref(muj)))
    out = {"TC": tc.value, "ry": sess.host(L.A_RY), "Qij": sess.host(L.A_QIJ), "uj": sess.host(L.A_UJ, squeeze=True)}
    L.check(sess.lib.lcx_direction_ns(sess.h, eps, C.byref(tang)))
    out["tangent"] = tang.value
    out["grad"] = sess.host(L.A_GRAD)
    out["update"] = sess.host(L.A_UPDATE)
    l0 = sess.launches()
    L.check(sess.lib.lcx_trial_ns(sess.h, eps, 1e-3, 0, C.byref(tc), C.byref(muj)))
    out["trial_launches"] = sess.launches() - l0
    out["TC_trial"] = tc.value
    out["ry_trial"] = sess.host(L.A_RY, 1)
    out["Qij_trial"] = sess.host(L.A_QIJ, 1)
    out["uj_trial"] = sess.host(L.A_UJ, 1, squeeze=True)
    sess.close()
    return out


@pytest.mark.parametrize("shape", [(300, 700, 130), (64, 1500, 200), (500, 129, 65), (200, 40, 3)])
@pytest.mark.parametrize("precision", ["fp64_split", "fp64_split5", "fast"])
def test_products_match_dmma(monkeypatch, shape, precision):
    """Same session state, the four products once through DMMA and once through the int8 engine: tile-edge shapes
    (m across the 64- and 128-wide factor tiles, n off every multiple of 128)."""
    import corex_oracle as oc
    N, n, m = shape
    x = oc.latent_factor_data(N, n, max(2, m // 4), seed=3, snr=1.0, snr_spread=0.3)
    xt = np.ascontiguousarray((x - x.mean(0)) / x.std(0), dtype=np.float64)
    rng = np.random.RandomState(5)
    w = np.ascontiguousarray(rng.randn(m, n) / (10.0 * np.sqrt(n)))
    a = _arrays_after_step(monkeypatch, "0", precision, xt, w, 0.36)
    b = _arrays_after_step(monkeypatch, "1", precision, xt, w, 0.36)
    tol = {"fp64_split": 1e-11, "fp64_split5": 1e-9, "fast": 1e-4}[precision]
    assert b.pop("trial_launches") > a.pop("trial_launches")  # the int8 path really ran (extra digit-slicing launches)
    for key in a:
        assert np.all(np.isfinite(a[key])), key
        assert_close(b[key], a[key], tol, "%s %s" % (key, precision))


def test_auto_rule_large_m(monkeypatch):
    """Without the override the int8 engine takes the products from m = 384 factors (n >= 2048) and gives what DMMA gives."""
    import corex_oracle as oc
    N, n, m = 96, 2100, 384
    x = oc.latent_factor_data(N, n, 40, seed=4, snr=1.0, snr_spread=0.3)
    xt = np.ascontiguousarray((x - x.mean(0)) / x.std(0), dtype=np.float64)
    w = np.ascontiguousarray(np.random.RandomState(6).randn(m, n) / (10.0 * np.sqrt(n)))
    a = _arrays_after_step(monkeypatch, "0", "fp64_split", xt, w, 0.0)
    b = _arrays_after_step(monkeypatch, None, "fp64_split", xt, w, 0.0)
    assert b.pop("trial_launches") > a.pop("trial_launches")
    for key in a:
        assert np.all(np.isfinite(a[key])), key
        assert_close(b[key], a[key], 1e-11, key)
    small = _arrays_after_step(monkeypatch, None, "fp64_split", xt[:, :300].copy(), w[:100, :300].copy(), 0.0)
    ref = _arrays_after_step(monkeypatch, "0", "fp64_split", xt[:, :300].copy(), w[:100, :300].copy(), 0.0)
    assert small["trial_launches"] == ref["trial_launches"]  # m = 100 stays on DMMA


def test_nonfinite_input_poisons_products(mm_i8):
    """A NaN in W must come out as NaN (like numpy in the reference), not be skipped by the column / row maxima."""
    import torch
    from linearcorex_b200 import _lib as L
    from linearcorex_b200.corex import _DeviceSession
    rng = np.random.RandomState(0)
    xt_np = rng.randn(100, 150)
    w = np.ascontiguousarray(rng.randn(4, 150) / 100.0)
    w[2, 17] = np.nan
    sess = _DeviceSession(L.PRECISION_FP64_SPLIT)
    ld = sess.lib.lcx_ld(150)
    xt = torch.zeros((100, ld), dtype=torch.float64, device="cuda")
    xt[:, :150] = torch.from_numpy(xt_np)
    sess.bind(xt, 100, 150, 4, None)
    L.check(sess.lib.lcx_set_w(sess.h, w.ctypes.data_as(C.c_void_p), 150))
    tc, muj = C.c_double(), C.c_double()
    L.check(sess.lib.lcx_moments_ns(sess.h, 0.0, 0, C.byref(tc), C.byref(muj)))
    assert not np.isfinite(tc.value)
    sess.close()
