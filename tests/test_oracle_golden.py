"""Pins oracle/corex_oracle.py against vectors produced by the unmodified reference
(tests/golden/*.npz, written by oracle/gen_golden.py).  CPU only."""
import numpy as np
import pytest

from conftest import load_golden, golden_moments
import corex_oracle as oc

FIT_CASES_F64 = ["readme_demo_f64", "big5_l0_f64", "big5_l1_f64", "test_data_f64", "adni_l1_f64", "adni_l2_f64",
                 "syn_400x300x10_f64", "syn_60x400x8_f64", "syn_400x300x10_noanneal_f64",
                 "syn_400x300x10_synergy_f64", "big5_syn_f64", "outliers_missing_f64", "outliers_f64",
                 "standard_missing_f64"]
FIT_CASES_NATIVE = ["readme_demo_native", "big5_l0_native", "big5_l0_cli", "test_data_cli",
                    "syn_400x300x10_native"]
RTOL = 1e-10  # same numpy expressions in the same order: differences are BLAS-threading noise only


def _fit_oracle(name, dtype):
    z, kw, x = load_golden(name)
    if name.startswith("readme_demo"):
        mdl = oc.OracleCorex(work_dtype=dtype, **kw)
        x = np.random.random((100, 50))  # README.md:49-51: drawn after the constructor seeded the RNG
        mdl.fit(x)
    else:
        mdl = oc.OracleCorex(work_dtype=dtype, **kw).fit(x)
    return z, mdl, x


def _check_fit(z, mdl, x, rtol):
    assert len(mdl.history["TC"]) == len(z["history_TC"])
    np.testing.assert_allclose(np.asarray(mdl.history["TC"], dtype=np.float64), z["history_TC"], rtol=rtol, atol=1e-12)
    assert mdl.ws.dtype == z["ws"].dtype
    np.testing.assert_allclose(mdl.ws, z["ws"], rtol=rtol, atol=1e-12)
    gm = golden_moments(z)
    assert set(gm) == set(mdl.moments)
    for key, val in gm.items():
        np.testing.assert_allclose(mdl.moments[key], val, rtol=max(rtol, 1e-8), atol=1e-10, err_msg=key)
    np.testing.assert_array_equal(mdl.clusters(), z["clusters"])
    np.testing.assert_allclose(mdl.theta[0], z["theta_mean"], rtol=rtol, atol=1e-14)
    np.testing.assert_allclose(mdl.theta[1], z["theta_std"], rtol=rtol, atol=1e-14)
    np.testing.assert_allclose(mdl.transform(x), z["transform"], rtol=1e-8, atol=1e-10)
    if "covariance" in z:
        np.testing.assert_allclose(mdl.get_covariance(), z["covariance"], rtol=1e-8, atol=1e-10)
    if "predict7" in z:
        np.testing.assert_allclose(mdl.predict(z["transform"][:7]), z["predict7"], rtol=1e-7, atol=1e-9)
    if "trials" in z and mdl.discourage_overlap:
        np.testing.assert_array_equal([t["trials"] for t in mdl.trace], z["trials"])
        np.testing.assert_array_equal([t["quick_fails"] for t in mdl.trace], z["quick_fails"])


@pytest.mark.parametrize("name", FIT_CASES_F64)
def test_oracle_fit_matches_reference_f64(name):
    z, mdl, x = _fit_oracle(name, np.float64)
    _check_fit(z, mdl, x, RTOL)


@pytest.mark.parametrize("name", FIT_CASES_NATIVE)
def test_oracle_fit_matches_reference_native(name):
    z, mdl, x = _fit_oracle(name, np.float32)
    _check_fit(z, mdl, x, 1e-5)  # float32 trajectories: summation-order noise is ~1e-6


def test_oracle_adni_layer0_cli():
    z, mdl, x = _fit_oracle("adni_l0_cli", np.float32)
    assert abs(float(mdl.tc) - float(z["m_TC"])) < 1e-6 * abs(float(z["m_TC"]))
    np.testing.assert_array_equal(mdl.clusters(), z["clusters"])


def test_known_answers():
    """Arithmetic-independent facts the data files pin."""
    z, _, _ = load_golden("test_data_f64")
    c = z["clusters"]
    assert c[0] == c[1] == c[2] and c[3] == c[4] and c[0] != c[3]
    z, _, _ = load_golden("big5_l0_f64")
    c = z["clusters"]
    assert len(set(c)) == 5 and all(c[i] == c[i % 5] for i in range(50))


@pytest.mark.parametrize("name", ["step_ns_400x300x10_f64", "step_ns_big5_f64", "step_ns_60x400x8_f64"])
def test_oracle_single_calls_ns(name):
    z, kw, _ = load_golden(name)
    xt, w = z["xt"], z["w"]
    for eps, tag in ((0.0, "e00_"), (0.36, "e36_")):
        for quick, p in ((True, "q_"), (False, "f_")):
            got = oc.moments_ns(xt, w, eps, quick=quick)
            want = golden_moments(z, tag + p)
            assert set(got) == set(want)
            for key in want:
                np.testing.assert_allclose(got[key], want[key], rtol=1e-9, atol=1e-12, err_msg=key)
        np.testing.assert_allclose(oc.sigma_times(xt, z[tag + "sig_u"], eps), z[tag + "sig"], rtol=1e-11, atol=1e-13)
        np.testing.assert_allclose(oc.norm_y(xt, w, eps), z[tag + "norm"], rtol=1e-12)
        rec = {}
        w2, m2 = oc.step_ns(xt, w, oc.moments_ns(xt, w, eps), eps, tol=1e-5, trace=rec)
        assert rec["trials"] == int(z[tag + "trials"])
        np.testing.assert_allclose(w2, z[tag + "w_next"], rtol=1e-9, atol=1e-12)
        for key, val in golden_moments(z, tag + "n_").items():
            np.testing.assert_allclose(m2[key], val, rtol=1e-8, atol=1e-11, err_msg=key)


def test_oracle_single_calls_syn():
    z, kw, _ = load_golden("step_syn_400x300x10_f64")
    xt, w = z["xt"], z["w"]
    got = oc.moments_syn(xt, w)
    want = golden_moments(z, "e00_f_")
    assert set(got) == set(want)
    for key in want:
        np.testing.assert_allclose(got[key], want[key], rtol=1e-9, atol=1e-12, err_msg=key)
    w2, m2 = oc.step_syn(xt, w, got, eta=0.1)
    np.testing.assert_allclose(w2, z["e00_w_next"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(m2["TC"], z["e00_n_TC"], rtol=1e-9)


def test_oracle_layers_match_reference_chain():
    """vis_corex.py:529-545 restated: layer k+1 fits on transform() of layer k."""
    z0, kw, x = load_golden("big5_l0_f64")
    z1, _, _ = load_golden("big5_l1_f64")
    models = oc.fit_layers(x, [5, 1], seed=0, work_dtype=np.float64)
    assert len(models) == 2
    np.testing.assert_allclose(float(models[0].tc), float(z0["m_TC"]), rtol=1e-10)
    np.testing.assert_allclose(float(models[1].tc), float(z1["m_TC"]), rtol=1e-8)
