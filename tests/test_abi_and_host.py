"""CPU-only checks: the C-ABI library loads and exports every symbol the header declares, the
workspace layout arithmetic works without a GPU, the product refuses to run without CUDA (no CPU
fallback), and the host-side helpers behave."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "lcx_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lcx_[a-z0-9_]+)\s*\(", src)) - {"lcx_allreduce_fn"})


def test_library_exports_every_declared_symbol():
    from linearcorex_b200 import _lib
    lib = _lib.load()
    declared = _header_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), "liblcx_b200.so lacks %s" % name
    assert sorted(_lib.SIGNATURES) == declared, set(_lib.SIGNATURES) ^ set(declared)
    assert lib.lcx_version() >= 100


def test_layout_arithmetic_without_gpu():
    from linearcorex_b200 import _lib
    lib = _lib.load()
    assert lib.lcx_ld(50) == 64 and lib.lcx_ld(10000) == 10000 and lib.lcx_ld(5) == 16
    assert lib.lcx_ldy(100) == 104 and lib.lcx_ldy(5) == 8
    small = lib.lcx_workspace_doubles(2000, 50, 5, 0)
    big = lib.lcx_workspace_doubles(100000, 10000, 100, 0)
    split = lib.lcx_workspace_doubles(100000, 10000, 100, 2)
    assert split > big + 6 * 100000 * 10000 // 8  # six int8 digit planes of X~
    assert lib.lcx_workspace_doubles(100000, 10000, 100, 3) < split  # five planes
    assert 0 < small < big
    # config 3 workspace: Y (100000 x 104) + ~20 m x n arrays + split-K partials, well under 1 GB
    assert big * 8 < 1.0e9
    assert lib.lcx_workspace_doubles(-1, 10, 2, 0) < 0
    assert lib.lcx_colstats_scratch_doubles(100000, 10000) > 0
    assert lib.lcx_project_scratch_doubles(100000, 100) >= 782 * 104


def test_layout_of_the_int8_product_route(monkeypatch):
    """The m x m x n products move to the int8 engine from m = 384 factors (split modes only): the workspace then also holds
    three m x n digit-plane operands, the planes of an m x m matrix and the per-column scale scratch."""
    from linearcorex_b200 import _lib
    lib = _lib.load()
    monkeypatch.delenv("LCX_MM_I8", raising=False)
    N, n, m = 1000, 50000, 500
    auto = lib.lcx_workspace_doubles(N, n, m, 2)
    monkeypatch.setenv("LCX_MM_I8", "0")
    off = lib.lcx_workspace_doubles(N, n, m, 2)
    monkeypatch.setenv("LCX_MM_I8", "1")
    on = lib.lcx_workspace_doubles(N, n, m, 2)
    assert auto == on > off
    ld8 = (n + 127) // 128 * 128
    assert on - off >= 3 * 6 * m * ld8 // 8 + 33 * n       # three operand plane sets + column scale scratch
    assert on - off < 4 * 6 * m * ld8 // 8 + 64 * n + 10 * m * m
    small_on = lib.lcx_workspace_doubles(100000, 10000, 100, 2)
    monkeypatch.delenv("LCX_MM_I8")
    assert lib.lcx_workspace_doubles(100000, 10000, 100, 2) < small_on   # m = 100 stays on DMMA unless forced
    assert lib.lcx_workspace_doubles(N, 1000, m, 2) == (monkeypatch.setenv("LCX_MM_I8", "0") or
                                                         lib.lcx_workspace_doubles(N, 1000, m, 2))  # n < 2048: DMMA
    monkeypatch.setenv("LCX_MM_I8", "1")
    assert lib.lcx_workspace_doubles(N, n, m, 0) == (monkeypatch.setenv("LCX_MM_I8", "0") or
                                                      lib.lcx_workspace_doubles(N, n, m, 0))      # DMMA mode ignores it


def _plan(lib, capfd, N, n, m, precision=2):
    """(K1, K2) plans of the split-integer contractions as printed by LCX_PLAN_DEBUG: dicts of splits / chunk / tail."""
    import re
    capfd.readouterr()
    assert lib.lcx_workspace_doubles(N, n, m, precision) > 0
    text = capfd.readouterr().err
    out = []
    for part in re.findall(r"splits (\d+) chunk (\d+) tail (\d+) tiles x (\d+) splits \(chunk (\d+)\)", text):
        out.append(dict(zip(("splits", "chunk", "tail_tiles", "tail_splits", "tail_chunk"), map(int, part))))
    assert len(out) == 2, text
    return out


def test_two_level_split_plan(monkeypatch, capfd):
    """host_session.cuh: plan_two_level.  With 12 500 samples per rank (config 3 on 8 GPUs) the 79 variable tiles of the
    second contraction run one full-length round on the 74 resident cluster pairs and the 5 left-over tiles are cut 14 ways;
    every chunk is a whole number of 64-deep K blocks, stays within the int32-exact contraction length, and the tail units
    fit one round.  No tail with an odd number of factor tiles (m = 30), in the DMMA mode, or when switched off."""
    from linearcorex_b200 import _lib
    lib = _lib.load()
    monkeypatch.setenv("LCX_PLAN_DEBUG", "1")
    monkeypatch.delenv("LCX_OZ_TAIL", raising=False)
    k1, k2 = _plan(lib, capfd, 12500, 10000, 100)
    assert (k2["splits"], k2["tail_tiles"], k2["tail_splits"]) == (1, 5, 14)
    assert (k1["splits"], k1["tail_tiles"], k1["tail_splits"]) == (1, 24, 3)
    kmax = (2 ** 31 // (127 * 127 * 6)) // 64 * 64
    for N, n, m in [(12500, 10000, 100), (25000, 10000, 100), (50000, 10000, 100), (100000, 10000, 100), (125000, 20000, 100),
                    (1000000, 20000, 100), (1000, 50000, 500), (9600, 9500, 120), (4000, 20000, 70), (777, 333, 65)]:
        for plan, K, m_tiles in zip(_plan(lib, capfd, N, n, m), (n, N), ((N + 127) // 128, (n + 127) // 128)):
            assert plan["chunk"] % 64 == 0 and plan["chunk"] <= kmax and plan["splits"] * plan["chunk"] >= K
            assert (plan["splits"] - 1) * plan["chunk"] < K
            if plan["tail_tiles"]:
                groups = (-(-m // 64) + 1) // 2
                assert plan["tail_chunk"] % 64 == 0 and plan["tail_splits"] >= 2
                assert plan["tail_tiles"] <= m_tiles
                assert groups * plan["tail_tiles"] * plan["tail_splits"] <= 74          # one round of the cluster pairs
                last = K - (plan["splits"] - 1) * plan["chunk"]                          # the K chunk the tail re-cuts
                assert (plan["tail_splits"] - 1) * plan["tail_chunk"] < last <= plan["tail_splits"] * plan["tail_chunk"]
                units = groups * m_tiles * plan["splits"]
                assert units - groups * plan["tail_tiles"] <= units // 74 * 74           # what is left fills whole rounds
    assert all(p["tail_tiles"] == 0 for p in _plan(lib, capfd, 125000, 20000, 30))       # one factor tile: no pairs
    monkeypatch.setenv("LCX_OZ_TAIL", "0")
    assert all(p["tail_tiles"] == 0 for p in _plan(lib, capfd, 12500, 10000, 100))
    uniform = _plan(lib, capfd, 12500, 10000, 100)
    assert uniform[1]["splits"] == 7 and uniform[0]["splits"] == 3                        # what round 1 ran at this shape


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from linearcorex_b200 import Corex, _lib
    with pytest.raises(_lib.LcxError):
        Corex(n_hidden=2, seed=0).fit(np.random.RandomState(0).randn(20, 6))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "linearcorex_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "corex_oracle" not in text and "/root/reference" not in text, f


def test_constructor_mirrors_reference_signature():
    from linearcorex_b200 import Corex
    c = Corex()
    assert (c.m, c.max_iter, c.tol, c.anneal, c.missing_values, c.discourage_overlap, c.gaussianize) == \
        (10, 10000, 1e-5, True, None, True, 'standard')
    assert c.ws.shape == (0, 0) and c.moments == {} and c.theta is None and c.eps == 0
    # seeding the global legacy RNG at construction (linearcorex.py:89)
    Corex(seed=3)
    a = np.random.randn(2)
    np.random.seed(3)
    assert np.array_equal(a, np.random.randn(2))
    assert Corex(eliminate_synergy=False).discourage_overlap is False
    with pytest.raises(ValueError):
        Corex(gaussianize="empirical")


def test_shard_rows_partition():
    from linearcorex_b200 import shard_rows
    for n, w in ((10, 3), (100000, 8), (7, 8), (1, 1)):
        spans = [shard_rows(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


def test_bench_data_is_row_shardable():
    import sys
    sys.path.insert(0, ROOT)
    import bench
    full = bench.make_rows(10000, 64, 4, 0, 10000, threads=2)
    part = bench.make_rows(10000, 64, 4, 3000, 9000, threads=2)
    assert np.array_equal(full[3000:9000], part)
    assert abs(full.std() - 1.0) < 0.02
    # planted structure: variables i and i+4 share a parent
    c = np.corrcoef(full[:, 0], full[:, 4])[0, 1]
    assert 0.4 < c < 0.6


def test_auto_precision_resolution():
    """'auto' picks the split-integer tcgen05 path except for launch-bound small problems (N n m < 3e7), where the DMMA
    path is quicker; explicit modes pass through."""
    from linearcorex_b200.corex import resolve_precision, AUTO_SPLIT_MIN_WORK
    assert resolve_precision("auto", 100, 50, 5) == "fp64"                 # README demo
    assert resolve_precision("auto", 2000, 50, 5) == "fp64"                # big5
    assert resolve_precision("auto", 566, 200, 30) == "fp64"               # adni
    assert resolve_precision("auto", 4000, 2000, 20) == "fp64_split"
    assert resolve_precision("auto", 100000, 10000, 100) == "fp64_split"   # config 3
    assert resolve_precision("auto", 1000000, 20000, 100) == "fp64_split"  # target (int overflow safe)
    assert resolve_precision("auto", 100000, 10000, 100, "none") == "fp64"  # un-normalised columns: one exponent is not enough
    assert resolve_precision("auto", 100000, 10000, 100, "outliers") == "fp64_split"
    assert AUTO_SPLIT_MIN_WORK == 3e7
    for mode in ("fp64", "fp64_split", "fp64_split5", "fp64_split7", "fast"):
        assert resolve_precision(mode, 10, 10, 1) == mode


def test_lazy_moments_is_a_dict_of_arrays():
    """LazyMoments (large models: m x n arrays fetched from the device on first read) must behave like the reference's plain
    dict: lookup, membership, get, iteration order, len, equality, copy and pickling."""
    import pickle
    import numpy as np
    from linearcorex_b200.corex import LazyMoments
    calls = []

    def fetch(name, value):
        def f():
            calls.append(name)
            return value
        return f
    rho, qij = np.arange(6.).reshape(2, 3), np.ones((2, 3))
    m = LazyMoments({"uj": np.zeros(2), "TC": 1.5}, {"rho": fetch("rho", rho), "Qij": fetch("Qij", qij)})
    assert "rho" in m and "TC" in m and "nope" not in m and len(m) == 4 and m.pending() == ["Qij", "rho"]
    assert m["TC"] == 1.5 and calls == []
    assert m["rho"] is rho and m["rho"] is rho and calls == ["rho"] and m.pending() == ["Qij"]
    assert m.get("additivity", 0) == 0 and m.get("Qij") is qij and calls == ["rho", "Qij"]
    m2 = LazyMoments({"TC": 2.0}, {"rho": fetch("rho2", rho)})
    assert sorted(m2) == ["TC", "rho"] and calls[-1] == "rho2"          # iteration materialises
    m3 = LazyMoments({"TC": 2.0}, {"rho": fetch("rho3", rho)})
    back = pickle.loads(pickle.dumps(m3))
    assert type(back) is dict and set(back) == {"TC", "rho"} and np.array_equal(back["rho"], rho)
    m4 = LazyMoments({"TC": 2.0}, {"rho": fetch("rho4", rho)})
    m4["rho"] = qij                                                     # overwriting drops the pending fetch
    assert m4["rho"] is qij and "rho4" not in calls and dict(m4.items()) == {"TC": 2.0, "rho": qij}
    try:
        m4["missing"]
        raise AssertionError("KeyError expected")
    except KeyError:
        pass
