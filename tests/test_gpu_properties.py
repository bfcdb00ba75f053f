"""Size-independent properties of the hot path at a shape well beyond the golden fixtures (20 000 x 3 000 x 40),
in both FP64-faithful modes, plus a fixed-budget parity run against the oracle at that shape.

Properties: rho(W) == _sig(x, W) exactly in exact arithmetic; _sig is linear; uj == diag(W rho^T);
max uj < 1 on every accepted iterate; TC never decreases inside an annealing stage (first Wolfe condition)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SHAPE = (20000, 3000, 40)


@pytest.fixture(scope="module")
def data():
    import corex_oracle as oc
    N, n, m = SHAPE
    x = oc.latent_factor_data(N, n, m, seed=2, snr=1.0, snr_spread=0.05).astype(np.float64)
    xt, theta, _ = oc.standardize(x, 'standard', None)
    rng = np.random.RandomState(1)
    w = rng.randn(m, n)
    w /= (10. * oc.norm_y(xt, w, 0.0))[:, None]
    return x, xt, w


def _session(xt, w, precision):
    import torch
    from linearcorex_b200 import _lib
    from linearcorex_b200.corex import _DeviceSession
    sess = _DeviceSession(_lib.PRECISIONS[precision])
    N, n = xt.shape
    ld = sess.lib.lcx_ld(n)
    xd = torch.zeros((N, ld), dtype=torch.float64, device="cuda")
    xd[:, :n] = torch.from_numpy(xt)
    sess.bind(xd, N, n, w.shape[0], None)
    _lib.check(sess.lib.lcx_set_w(sess.h, np.ascontiguousarray(w).ctypes.data_as(C.c_void_p), n))
    return sess, _lib, torch


def _sig(sess, L, torch, u, eps):
    m, n = u.shape
    ld = sess.lib.lcx_ld(n)
    ud = torch.zeros((m, ld), dtype=torch.float64, device="cuda")
    ud[:, :n] = torch.from_numpy(np.ascontiguousarray(u))
    od = torch.zeros_like(ud)
    L.check(sess.lib.lcx_sig(sess.h, ud.data_ptr(), eps, od.data_ptr()))
    return od[:, :n].cpu().numpy()


@pytest.mark.parametrize("precision,tol", [("fp64", 1e-12), ("fp64_split", 2e-10)])
def test_sig_identities(data, precision, tol):
    x, xt, w = data
    sess, L, torch = _session(xt, w, precision)
    eps = 0.36
    tc, muj = C.c_double(), C.c_double()
    L.check(sess.lib.lcx_moments_ns(sess.h, eps, 0, C.byref(tc), C.byref(muj)))
    rho = sess.host(L.A_RHO)
    uj = sess.host(L.A_UJ, squeeze=True)
    sig_w = _sig(sess, L, torch, w, eps)
    scale = np.abs(rho).max()
    assert np.abs(rho - sig_w).max() <= tol * scale                       # rho(W) == _sig(x, W)
    assert np.abs(uj - (w * rho).sum(1)).max() <= max(tol, 1e-12) * uj.max()  # uj == diag(W rho^T)
    rng = np.random.RandomState(3)
    u, v = rng.randn(*w.shape) * 0.01, rng.randn(*w.shape) * 0.3
    a, b = 0.7, -1.9
    lhs = _sig(sess, L, torch, a * u + b * v, eps)
    rhs = a * _sig(sess, L, torch, u, eps) + b * _sig(sess, L, torch, v, eps)
    assert np.abs(lhs - rhs).max() <= 4 * tol * np.abs(lhs).max()         # linearity
    # against numpy float64 directly
    want = (1 - eps ** 2) * (xt.T @ (xt @ w.T)).T / xt.shape[0] + eps ** 2 * w
    assert np.abs(sig_w - want).max() <= max(tol, 1e-12) * np.abs(want).max()
    sess.close()


@pytest.mark.parametrize("precision", ["fp64", "fp64_split"])
def test_fixed_budget_fit_matches_oracle_and_invariants(data, precision):
    import corex_oracle as oc
    from linearcorex_b200 import Corex
    x, xt, w = data
    N, n, m = SHAPE
    kw = dict(n_hidden=m, seed=0, max_iter=4, tol=1e-12)
    mdl = Corex(precision=precision, **kw).fit(x)
    ref = oc.OracleCorex(work_dtype=np.float64, **kw).fit(x)
    assert len(mdl.history["TC"]) == len(ref.history["TC"]) == 28
    err_w = np.abs(mdl.ws - ref.ws).max() / np.abs(ref.ws).max()
    err_t = np.abs(mdl.tcs - ref.tcs).max() / np.abs(ref.tcs).max()
    assert err_w < 1e-9 and err_t < 1e-9, (err_w, err_t)
    np.testing.assert_array_equal(mdl.clusters(), ref.clusters())
    np.testing.assert_array_equal([t["trials"] for t in mdl.trace], [t["trials"] for t in ref.trace])
    assert mdl.moments["uj"].max() < 1.0
    # TC is non-decreasing within each annealing stage (Wolfe sufficient increase)
    tcs = np.asarray(mdl.history["TC"]).reshape(7, 4)
    assert (np.diff(tcs, axis=1) >= -1e-9 * np.abs(tcs[:, 1:])).all()


def test_config4_like_small_sample_path():
    """n >> N (under-sampled, gaussianize='outliers', m > 128): exercises the split-over-variables first contraction,
    multi-tile factors and the m x m inverse at m = 160."""
    import corex_oracle as oc
    from linearcorex_b200 import Corex
    x = oc.latent_factor_data(300, 6000, 160, seed=4, snr=3.0, snr_spread=0.01).astype(np.float64)
    kw = dict(n_hidden=160, seed=0, max_iter=2, tol=1e-12, gaussianize="outliers")
    ref = oc.OracleCorex(work_dtype=np.float64, **kw).fit(x)
    for precision in ("fp64", "fp64_split"):
        mdl = Corex(precision=precision, **kw).fit(x)
        assert len(mdl.history["TC"]) == len(ref.history["TC"])
        err_w = np.abs(mdl.ws - ref.ws).max() / np.abs(ref.ws).max()
        err_tc = abs(mdl.tc - ref.tc) / abs(ref.tc)
        assert err_w < 1e-8 and err_tc < 1e-9, (precision, err_w, err_tc)
        np.testing.assert_allclose(mdl.moments["X_i Z_j"], ref.moments["X_i Z_j"], rtol=0, atol=1e-7 * np.abs(ref.moments["X_i Z_j"]).max())


@pytest.mark.parametrize("precision", ["fp64", "fp64_split", "fast"])
def test_run_to_run_bit_identical(precision):
    """No float atomics anywhere: split-K partials and per-CTA reductions are combined in fixed order, so repeated fits of
    the same data are bit-identical (shape chosen so that both contractions run split-K over several waves of CTAs)."""
    import corex_oracle as oc
    from linearcorex_b200 import Corex
    x = oc.latent_factor_data(30000, 1500, 40, seed=4, snr=1.0, snr_spread=0.5)
    fits = [Corex(n_hidden=40, seed=3, max_iter=4, tol=1e-12, precision=precision).fit(x) for _ in range(3)]
    for other in fits[1:]:
        assert np.array_equal(other.ws, fits[0].ws)
        assert other.history["TC"] == fits[0].history["TC"]
        for key, val in fits[0].moments.items():
            assert np.array_equal(np.asarray(other.moments[key]), np.asarray(val)), key


def test_lazy_moments_export_equals_eager(monkeypatch):
    """Large models keep the m x n arrays of `moments` on the device until a key is read (Corex.LAZY_MOMENTS_BYTES); forced on
    for a golden case, every key -- and what is derived from them -- must equal the eager export bit for bit."""
    import pickle
    from conftest import load_golden
    from linearcorex_b200 import Corex
    from linearcorex_b200.corex import LazyMoments
    z, kw, x = load_golden("syn_400x300x10_f64")
    eager = Corex(precision="fp64_split", **kw).fit(x)
    monkeypatch.setattr(Corex, "LAZY_MOMENTS_BYTES", 0)
    lazy = Corex(precision="fp64_split", **kw).fit(x)
    assert isinstance(lazy.moments, LazyMoments) and not isinstance(eager.moments, LazyMoments)
    assert {"rho", "Qij", "MI", "X_i Z_j"} <= set(lazy.moments.pending())
    assert lazy.tc == eager.tc and np.array_equal(lazy.tcs, eager.tcs) and "rho" in lazy.moments.pending()
    np.testing.assert_array_equal(lazy.get_covariance(), eager.get_covariance())   # reads rhoinvrho / Si only
    assert "rho" in lazy.moments.pending() and "rhoinvrho" not in lazy.moments.pending()
    Corex(precision="fp64_split", **kw).fit(x[::-1].copy())   # a later fit reuses device memory: the snapshots must not move
    back = pickle.loads(pickle.dumps(lazy))
    assert type(back.moments) is dict and list(back.moments) == list(eager.moments)
    for key, val in eager.moments.items():
        np.testing.assert_array_equal(np.asarray(lazy.moments[key]), np.asarray(val), err_msg=key)
        np.testing.assert_array_equal(np.asarray(back.moments[key]), np.asarray(val), err_msg=key)
