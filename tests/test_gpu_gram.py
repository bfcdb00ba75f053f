"""GPU parity of the Gram route (`Corex(algorithm='gram')`, lcx_gram_build / lcx_bind_gram, csrc/host_gram.cuh).

The fit depends on the data only through X~^T X~ / N (linearcorex.py:196-213 is u -> (X~^T X~ / N) u^T; :248 is
a_j^T (X~^T X~ / N) a_j), so for N >= n the matrix is formed once -- as exact int8 digit products on tcgen05 -- and every
pass pair of the loop becomes one n x n x m product.  Same bar as the streaming route: the goldens written by the unmodified
reference (float64 path) to 1e-9 on W / moments / TCs with identical iteration counts and bit-exact clusters; the matrix
itself against numpy to 1e-13."""
import ctypes as C

import numpy as np
import pytest

from conftest import load_golden
from test_gpu_parity import _check_fit, _fit, assert_close, RTOL

pytestmark = pytest.mark.gpu


def _gram_matrix(xt_np, precision, block_cols):
    import torch
    from linearcorex_b200 import _lib
    from linearcorex_b200.corex import _DeviceSession
    sess = _DeviceSession(_lib.PRECISIONS[precision])
    lib = sess.lib
    N, n = xt_np.shape
    ld = lib.lcx_ld(n)
    xt = torch.zeros((N, ld), dtype=torch.float64, device="cuda")
    xt[:, :n] = torch.from_numpy(np.ascontiguousarray(xt_np, dtype=np.float64))
    sess.bind(xt, N, n, 4, None)
    g = torch.full((n, ld), float("nan"), dtype=torch.float64, device="cuda")
    nscr = lib.lcx_gram_scratch_doubles(sess.h, block_cols)
    assert nscr > 0
    scratch = torch.empty(nscr, dtype=torch.float64, device="cuda")
    _lib.check(lib.lcx_gram_build(sess.h, g.data_ptr(), ld, block_cols, scratch.data_ptr(), nscr), "lcx_gram_build")
    out = g[:, :n].cpu().numpy()
    sess.close()
    return out


@pytest.mark.parametrize("shape,block", [((700, 300), 128), ((700, 300), 256), ((129, 127), 128), ((5000, 520), 512),
                                         ((300, 130), 128), ((40, 200), 128), ((23000, 257), 128)])
@pytest.mark.parametrize("precision", ["fp64_split", "fp64_split7"])
def test_gram_matrix_matches_numpy(shape, block, precision):
    """X~^T X~ / N over tile edges, several column blocks, N < n, and more samples than one int32-exact split (23 000)."""
    rng = np.random.RandomState(sum(shape) + block)
    x = rng.randn(*shape)
    x[:, 1::3] += 0.7 * x[:, :1]          # correlated columns
    x = (x - x.mean(0)) / x.std(0)
    got = _gram_matrix(x, precision, block)
    want = x.T.dot(x) / shape[0]
    assert np.isfinite(got).all()
    np.testing.assert_array_equal(got, got.T)   # mirrored exactly
    # 6 digits: the digit planes keep 48 bits below max |X~| -- an absolute perturbation of 2^E 254^-6 per entry
    assert np.abs(got - want).max() <= (3e-13 if precision == "fp64_split" else 3e-15) * np.abs(want).max()


GRAM_CASES = ["readme_demo_f64", "big5_l0_f64", "syn_400x300x10_f64", "syn_400x300x10_noanneal_f64", "outliers_missing_f64",
              "outliers_f64", "standard_missing_f64", "adni_l1_f64", "big5_l1_f64", "adni_l2_f64", "syn_4000x2000x20_f64"]


@pytest.mark.parametrize("name", GRAM_CASES)
def test_full_fit_gram(name):
    z, mdl, x = _fit(name, precision="fp64_split", algorithm="gram")
    assert mdl.algorithm_used == "gram"
    _check_fit(z, mdl, x, RTOL)


@pytest.mark.parametrize("name", ["syn_400x300x10_f64", "big5_l0_f64"])
def test_full_fit_gram_exact_trials(name):
    """The reference's control flow literally (one product with the matrix per trial): trial counts equal the reference's."""
    z, mdl, x = _fit(name, precision="fp64_split", algorithm="gram", exact_trials=True)
    _check_fit(z, mdl, x, RTOL)


@pytest.mark.parametrize("name", ["syn_400x300x10_synergy_f64", "big5_syn_f64"])
def test_full_fit_gram_synergy(name):
    z, mdl, x = _fit(name, precision="fp64_split", algorithm="gram")
    _check_fit(z, mdl, x, RTOL)


def test_gram_adni_layer0_long_trajectory():
    z, mdl, x = _fit("adni_l0_f64", precision="fp64_split", algorithm="gram")
    assert len(mdl.history["TC"]) == len(z["history_TC"])
    assert_close(mdl.ws, z["ws"], RTOL, "ws")
    assert_close(mdl.tcs, z["m_TCs"], RTOL, "TCs")
    np.testing.assert_array_equal(mdl.clusters(), z["clusters"])


def test_gram_streamed_preparation():
    """Row-block streaming of the raw input feeds the same digit planes, so the matrix -- and the fit -- are unchanged."""
    z, mdl, x = _fit("syn_400x300x10_f64", precision="fp64_split", algorithm="gram", stream_rows=96)
    _check_fit(z, mdl, x, RTOL)


def test_gram_auto_rule_and_stream_agreement():
    """'auto' picks the Gram route for N >= n once the problem is bound by the passes over X; both routes agree to 1e-9."""
    import corex_oracle as oc
    from linearcorex_b200 import Corex
    x = oc.latent_factor_data(30000, 1500, 24, seed=3, snr=1.0, snr_spread=0.05)
    kw = dict(n_hidden=24, seed=0, max_iter=4, tol=1e-12)
    auto = Corex(precision="fp64_split", **kw).fit(x)
    assert auto.algorithm_used == "gram"                       # 30 000 x 1 500 x 24 = 1.08e9
    stream = Corex(precision="fp64_split", algorithm="stream", **kw).fit(x)
    assert stream.algorithm_used == "stream"
    assert len(auto.history["TC"]) == len(stream.history["TC"])
    assert [t["trials"] for t in auto.trace] == [t["trials"] for t in stream.trace]
    assert_close(auto.ws, stream.ws, RTOL, "ws")
    assert_close(auto.tcs, stream.tcs, RTOL, "TCs")
    assert_close(np.asarray(auto.history["TC"]), np.asarray(stream.history["TC"]), RTOL, "history")
    np.testing.assert_array_equal(auto.clusters(), stream.clusters())
    small = Corex(precision="fp64_split", n_hidden=5, seed=0, max_iter=2).fit(x[:400, :300])
    assert small.algorithm_used == "stream"                    # launch-bound problem: nothing to gain
    wide = Corex(precision="fp64_split", n_hidden=40, seed=0, max_iter=1, tol=1e-12).fit(
        np.random.RandomState(0).randn(1000, 30000))
    assert wide.algorithm_used == "stream"                     # n >> N: the reference's own formulation
    with pytest.raises(ValueError):
        Corex(precision="fp64", algorithm="gram", n_hidden=3).fit(x[:200, :50])


def test_gram_run_to_run_bit_identical():
    from linearcorex_b200 import Corex
    _, kw, x = load_golden("syn_400x300x10_f64")
    a = Corex(precision="fp64_split", algorithm="gram", **kw).fit(x)
    b = Corex(precision="fp64_split", algorithm="gram", **kw).fit(x)
    np.testing.assert_array_equal(a.ws, b.ws)
    assert a.history["TC"] == b.history["TC"]


def test_gram_warm_start_and_refit_on_the_same_object():
    """A second fit on a fitted model (warm start, :114) re-binds the data session and builds a fresh matrix; both routes must
    agree on the refit, and `transform` / `get_covariance` / pickling work from a Gram-route model like from any other."""
    import pickle
    from linearcorex_b200 import Corex
    _, kw, x = load_golden("syn_400x300x10_f64")
    kw = dict(kw, max_iter=5, tol=1e-12)
    g = Corex(precision="fp64_split", algorithm="gram", **kw).fit(x)
    s = Corex(precision="fp64_split", algorithm="stream", **kw).fit(x)
    assert_close(g.ws, s.ws, RTOL, "first fit")
    g.fit(x)
    s.fit(x)
    assert g.algorithm_used == "gram" and s.algorithm_used == "stream"
    assert len(g.history["TC"]) == len(s.history["TC"])
    assert_close(g.ws, s.ws, RTOL, "warm-started refit")
    assert_close(g.transform(x[:50]), s.transform(x[:50]), RTOL, "transform")
    assert_close(g.get_covariance(), s.get_covariance(), RTOL, "covariance")
    back = pickle.loads(pickle.dumps(g))
    assert back.algorithm_used == "gram"
    assert_close(back.get_covariance(), s.get_covariance(), RTOL, "covariance after unpickle")
    assert_close(back.transform(x[:50]), s.transform(x[:50]), RTOL, "transform after unpickle")
