"""Oracle parity at the geometries bench.py times (BASELINE.json configs[2], configs[3], configs[4]) -- the tile edges,
wave counts and automatic kernel routes of the benchmarked shapes, at row counts the CPU oracle finishes in seconds.

  config 3   100 000 x 10 000 x 100: n = 10 000 variables and m = 100 factors are kept (the 64 + 48 factor tiles in
             lock-step, 79 variable tiles, the split-K plan of the second contraction); 8 192 sample rows.
  config 4   1 000 x 50 000 x 500, gaussianize='outliers': m = 500 routes the four m x m x n products of an iteration to
             the int8 engine automatically (m >= 384) and the details path through the cooperative LU; 8 192 variables.
  config 5   layers 30,5,1 on 1M x 20k: a two-layer fit_layers run at m = 30 on 50 000 x 2 000, layer 1 fed by the
             device-resident transform of layer 0.

All in the bench's FP64-faithful mode (fp64_split) at a fixed iteration budget, against oracle/corex_oracle.py in float64,
to the north star's 1e-9 (W, TC, per-factor TCs, every moments key; clusters bit-exact)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-9


def _rel(a, b, floor=0.0):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), floor, 1e-300)


def _compare(mdl, ref, tol=RTOL):
    assert len(mdl.history["TC"]) == len(ref.history["TC"])
    assert [t["trials"] for t in mdl.trace] == [t["trials"] for t in ref.trace]
    errs = {"ws": _rel(mdl.ws, ref.ws), "TC": _rel(mdl.tc, ref.tc), "TCs": _rel(mdl.tcs, ref.tcs),
            "history": _rel(mdl.history["TC"], ref.history["TC"])}
    tc_scale = abs(float(ref.tc))
    assert set(mdl.moments) == set(ref.moments)
    for key, val in ref.moments.items():
        errs[key] = _rel(mdl.moments[key], val, floor=tc_scale if key in ("additivity", "TC_direct", "TC_no_overlap") else 0.0)
    bad = {k: v for k, v in errs.items() if not v <= tol}
    assert not bad, "relative errors above %.0e: %s" % (tol, bad)
    np.testing.assert_array_equal(mdl.clusters(), ref.clusters())
    return errs


def test_config3_geometry_8192x10000x100():
    import corex_oracle as oc
    from linearcorex_b200 import Corex
    x = oc.latent_factor_data(8192, 10000, 100, seed=0, snr=1.0, snr_spread=0.02)
    kw = dict(n_hidden=100, seed=0, max_iter=2, tol=1e-12)
    mdl = Corex(precision="fp64_split", **kw).fit(x)
    assert mdl._sess.lib.lcx_ld(10000) == 10000 and mdl.precision_used == "fp64_split"
    ref = oc.OracleCorex(work_dtype=np.float64, **kw).fit(x)
    _compare(mdl, ref)
    assert _rel(mdl.transform(x[:300]), ref.transform(x[:300])) <= RTOL


def test_config3_geometry_gram_route_12288x10000x100():
    """The Gram route at the benchmarked n = 10 000, m = 100: 79 row tiles of the matrix, its split-K product, the column-block
    build with a partial last block, upper triangle + mirror -- N >= n so `algorithm='auto'` takes it."""
    import corex_oracle as oc
    from linearcorex_b200 import Corex
    x = oc.latent_factor_data(12288, 10000, 100, seed=4, snr=1.0, snr_spread=0.02)
    kw = dict(n_hidden=100, seed=0, max_iter=1, tol=1e-12)
    mdl = Corex(precision="fp64_split", **kw).fit(x)
    assert mdl.algorithm_used == "gram"
    ref = oc.OracleCorex(work_dtype=np.float64, **kw).fit(x)
    _compare(mdl, ref)


def test_config4_geometry_1000x8192x500_outliers():
    import corex_oracle as oc
    from linearcorex_b200 import Corex
    x = oc.latent_factor_data(1000, 8192, 500, seed=1, snr=2.0, snr_spread=0.004)
    x[:, ::9] = np.sign(x[:, ::9]) * np.abs(x[:, ::9]) ** 3  # long tails for g()
    kw = dict(n_hidden=500, seed=0, max_iter=2, tol=1e-12, gaussianize="outliers")
    mdl = Corex(precision="fp64_split", **kw).fit(x)
    ref = oc.OracleCorex(work_dtype=np.float64, **kw).fit(x)
    _compare(mdl, ref)


def test_config5_two_layers_m30_50000x2000():
    import corex_oracle as oc
    from linearcorex_b200 import fit_layers
    x = oc.latent_factor_data(50000, 2000, 30, seed=2, snr=1.0, snr_spread=0.05)
    kw = dict(seed=0, max_iter=3, tol=1e-12)
    ours = fit_layers(x, [30, 1], precision="fp64_split", **kw)
    ref = oc.fit_layers(x, [30, 1], work_dtype=np.float64, **kw)
    assert [m.m for m in ours] == [30, 1]
    _compare(ours[0], ref[0])
    # layer 1 sees layer 0's Y (device-resident here): its input already carries layer 0's 1e-13, amplified by a one-factor fit
    assert _rel(ours[1].ws, ref[1].ws) <= RTOL and _rel(ours[1].tc, ref[1].tc) <= RTOL
