"""Edge cases of the fit path against the oracle (oracle/corex_oracle.py, pinned to the reference by
tests/test_oracle_golden.py): degenerate shapes, odd input containers and dtypes, badly scaled or constant columns,
inputs the reference turns into NaN.  Fixed small iteration budgets so that both sides stop at the same iterate."""
import numpy as np
import pytest

from test_gpu_parity import assert_close

pytestmark = pytest.mark.gpu

BUDGET = dict(max_iter=5, tol=1e-12, seed=1)


def _pair(x, precision, oracle_x=None, **kw):
    import corex_oracle as oc
    from linearcorex_b200 import Corex
    kw = dict(dict(BUDGET, n_hidden=2), **kw)
    ref = oc.OracleCorex(work_dtype=np.float64, **kw).fit(np.array(x if oracle_x is None else oracle_x, dtype=np.float64))
    mdl = Corex(precision=precision, **kw).fit(x)
    return mdl, ref


def _same_fit(mdl, ref, tol, floor=0.0):
    assert len(mdl.history["TC"]) == len(ref.history["TC"])
    assert mdl.ws.shape == ref.ws.shape
    assert_close(mdl.ws, ref.ws, tol, "ws")
    assert_close(mdl.tc, ref.tc, tol, "TC", floor=floor)
    assert_close(mdl.tcs, ref.tcs, tol, "TCs", floor=max(floor, abs(float(ref.tc))))
    np.testing.assert_array_equal(mdl.clusters(), ref.clusters())


def _data(shape, seed=3):
    return np.random.RandomState(seed).randn(*shape)


@pytest.mark.parametrize("precision", ["fp64", "fp64_split"])
@pytest.mark.parametrize("shape,m", [((3, 2), 1), ((6, 3), 5), ((40, 2), 1), ((9, 7), 7), ((130, 129), 3), ((257, 17), 17)])
def test_degenerate_shapes(shape, m, precision):
    """Fewer samples than variables, more factors than variables, sizes straddling the tile edges.  (A single variable
    is left out on purpose: there the stage rescale factor of :130-133 is exactly 1 in exact arithmetic, so
    `0.001 * floor(1000 * a)` is 0.999 or 1.0 depending on the last bit of uj -- not a parity target.)"""
    x = _data(shape)
    mdl, ref = _pair(x, precision, n_hidden=m)
    _same_fit(mdl, ref, 1e-9, floor=1e-6)


def test_single_variable_runs():
    from linearcorex_b200 import Corex
    mdl = Corex(n_hidden=1, **BUDGET).fit(_data((40, 1)))
    assert mdl.ws.shape == (1, 1) and abs(abs(mdl.ws[0, 0]) - 0.1) < 1e-3 and abs(mdl.tc) < 1e-12


@pytest.mark.parametrize("precision", ["fp64", "fp64_split"])
def test_constant_column(precision):
    """sigma is clipped at 1e-10 (:415): a constant column standardises to zeros and must not poison the rest."""
    x = _data((50, 6))
    x[:, 2] = 7.0
    mdl, ref = _pair(x, precision)
    _same_fit(mdl, ref, 1e-9)
    assert np.isfinite(mdl.ws).all()


@pytest.mark.parametrize("precision", ["fp64", "fp64_split"])
def test_column_scales_twenty_four_orders_apart(precision):
    x = _data((50, 6))
    x[:, 1] *= 1e12
    x[:, 3] *= 1e-12
    mdl, ref = _pair(x, precision)
    _same_fit(mdl, ref, 1e-9)
    assert_close(mdl.theta[1], ref.theta[1], 1e-12, "sigma")


@pytest.mark.parametrize("gaussianize", ["standard", "outliers"])
def test_single_huge_outlier_fp64(gaussianize):
    """One entry at 1e6 in unit-variance data: |x~| reaches ~14 in 'standard' mode.  The DMMA mode has no exponent
    sharing and keeps the 1e-9 bar; the split modes document this as their limit -- see the next test."""
    x = _data((200, 8))
    x[5, 3] = 1e6
    mdl, ref = _pair(x, "fp64", gaussianize=gaussianize)
    _same_fit(mdl, ref, 1e-9)


@pytest.mark.parametrize("gaussianize", ["standard", "outliers"])
def test_single_huge_outlier_split(gaussianize):
    """Pure noise plus one outlier is an ill-conditioned fit: it amplifies the 48-bit mode's perturbation (one exponent for
    all of X~, |x~| up to 14 here) to 1e-7 under 'standard'.  The 56-bit mode is back at 1e-9."""
    x = _data((200, 8))
    x[5, 3] = 1e6
    mdl, ref = _pair(x, "fp64_split", gaussianize=gaussianize)
    _same_fit(mdl, ref, 1e-9 if gaussianize == "outliers" else 1e-7)
    mdl, ref = _pair(x, "fp64_split7", gaussianize=gaussianize)
    _same_fit(mdl, ref, 1e-9)


@pytest.mark.parametrize("precision", ["fp64", "fp64_split"])
def test_input_containers_and_dtypes(precision):
    """np.asarray semantics of :108: lists, integer arrays, Fortran order, strided views, float32, host and device tensors."""
    import torch
    base = np.random.RandomState(5).randint(0, 7, size=(80, 6)).astype(np.float64)
    base += 0.25 * np.random.RandomState(6).randint(0, 4, size=(80, 6))  # exactly representable in float32
    wide = np.zeros((160, 12))
    wide[::2, ::2] = base
    variants = {
        "list": base.tolist(),
        "fortran": np.asfortranarray(base),
        "strided": wide[::2, ::2],
        "float32": base.astype(np.float32),
        "host tensor": torch.from_numpy(base.copy()),
        "pinned float32 tensor": torch.from_numpy(base.astype(np.float32)).pin_memory(),
        "cuda float64 tensor": torch.from_numpy(base.copy()).cuda(),
        "cuda float32 tensor": torch.from_numpy(base.astype(np.float32)).cuda(),
    }
    mdl0, ref = _pair(base, precision)
    _same_fit(mdl0, ref, 1e-9)
    for name, v in variants.items():
        mdl, _ = _pair(v, precision, oracle_x=base)
        assert np.array_equal(mdl.ws, mdl0.ws), name  # same values in -> bit-identical fit, whatever the container
    ints = np.random.RandomState(7).randint(0, 5, size=(40, 6))
    mdl, ref = _pair(ints, precision)
    _same_fit(mdl, ref, 1e-9)


def test_pandas_frame_input():
    pd = pytest.importorskip("pandas")
    x = _data((60, 5))
    mdl, ref = _pair(pd.DataFrame(x, columns=list("abcde")), "fp64_split", oracle_x=x)
    _same_fit(mdl, ref, 1e-9)


@pytest.mark.parametrize("precision", ["fp64", "fp64_split"])
def test_non_finite_input_propagates_like_the_reference(precision, capsys):
    """An inf entry (or a column with no observed value) makes TC NaN in the reference, which prints and carries on
    (:144-149); no exception, no hang, NaN out."""
    from linearcorex_b200 import Corex
    x = _data((30, 5))
    x[3, 2] = np.inf
    mdl = Corex(n_hidden=2, precision=precision, **BUDGET).fit(x)
    assert not np.isfinite(mdl.tc)
    if precision != "fp64":  # streamed preparation derives the X~ scale on the host: same outcome
        mdl = Corex(n_hidden=2, precision=precision, stream_rows=16, **BUDGET).fit(x)
        assert not np.isfinite(mdl.tc)
    x = _data((30, 5))
    x[:, 0] = np.nan
    mdl = Corex(n_hidden=2, precision=precision, missing_values=np.nan, **BUDGET).fit(x)
    assert not np.isfinite(mdl.tc)
    capsys.readouterr()


def test_bad_arguments_raise():
    from linearcorex_b200 import Corex
    with pytest.raises((ValueError, IndexError)):
        Corex(n_hidden=2, **BUDGET).fit(_data((30,)))  # 1-D input: the reference fails unpacking x.shape (:110)
    with pytest.raises(ValueError):
        Corex(n_hidden=2, precision="fp16")
    with pytest.raises(ValueError):
        Corex(n_hidden=2, gaussianize="empirical")
    mdl = Corex(n_hidden=2, **BUDGET).fit(_data((30, 5)))
    with pytest.raises(AssertionError):
        mdl.transform(_data((4, 6)))  # :391


@pytest.mark.parametrize("precision", ["fp64", "fp64_split"])
def test_transform_of_one_row_and_of_unseen_rows(precision):
    x = _data((120, 9))
    mdl, ref = _pair(x, precision, n_hidden=3)
    fresh = _data((33, 9), seed=11)
    assert_close(mdl.transform(fresh), ref.transform(fresh), 1e-9, "transform unseen")
    assert_close(mdl.transform(fresh[:1]), ref.transform(fresh[:1]), 1e-9, "transform one row")
    lab = mdl.transform(fresh, details=True)
    assert lab[0].shape == (33, 3)


@pytest.mark.parametrize("overlap", [True, False])
def test_verbose_history_keys(overlap, capsys):
    """update_records (:166-175): with verbose set, `history` also carries per-iteration additivity and TCs -- zeros from
    the quick non-synergy moments (which have neither key), real values from the synergy moments."""
    from linearcorex_b200 import Corex
    x = _data((80, 7))
    mdl = Corex(n_hidden=2, verbose=1, discourage_overlap=overlap, **BUDGET).fit(x)
    capsys.readouterr()
    n_it = len(mdl.history["TC"])
    assert n_it > 0 and len(mdl.history["additivity"]) == n_it and len(mdl.history["TCs"]) == n_it
    assert all(np.shape(t) == (2,) for t in mdl.history["TCs"])
    if overlap:
        assert not np.any(mdl.history["additivity"]) and not np.any(mdl.history["TCs"])
    else:
        assert np.all(np.isfinite(mdl.history["TCs"])) and np.any(mdl.history["TCs"])
    quiet = Corex(n_hidden=2, discourage_overlap=overlap, **BUDGET).fit(x)
    assert sorted(quiet.history) == ["TC"] and np.array_equal(quiet.ws, mdl.ws)


def test_default_precision_is_auto():
    """Default `precision='auto'`: DMMA for launch-bound small problems, the split-integer path otherwise; a refit of the
    same object on a larger problem switches sessions."""
    import corex_oracle as oc
    from linearcorex_b200 import Corex
    small = _data((200, 12))
    mdl = Corex(n_hidden=3, **BUDGET)
    assert mdl.precision == "auto" and mdl.precision_used is None
    mdl.fit(small)
    assert mdl.precision_used == "fp64" and mdl._sess.xt is not None
    ref = oc.OracleCorex(work_dtype=np.float64, n_hidden=3, **BUDGET).fit(small)
    _same_fit(mdl, ref, 1e-9)
    big = oc.latent_factor_data(4000, 2000, 20, seed=1, snr=1.0, snr_spread=0.2)
    mdl2 = Corex(n_hidden=20, **dict(BUDGET, max_iter=2)).fit(big)
    assert mdl2.precision_used == "fp64_split" and mdl2._sess.xt is None
    mdl.m = 20
    mdl.ws = np.zeros((0, 0))
    mdl.fit(big)  # same object, new shape: 'auto' resolves again and the session is rebuilt
    assert mdl.precision_used == "fp64_split"
