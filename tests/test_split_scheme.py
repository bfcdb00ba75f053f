"""The split-integer (int8 digit plane) scheme of csrc/ozaki_i8.cuh, restated in numpy and checked on CPU:
digit ranges, exactness of the int32 group sums under the enforced contraction-length cap, and the accuracy of the
recombined product against float64.  (The CUDA kernels are checked against the same float64 results in the GPU tests.)"""
import numpy as np
import pytest


def pow2_above(amax):
    if not (amax > 0):
        return 1.0
    _, e = np.frexp(amax)
    return float(np.ldexp(1.0, int(e) + 1))


def split_digits(x, scale, S, R):
    """v = x / scale in (-0.5, 0.5); repeat t = R v, d = rint(t), v = t - d  (split_digits<S> in ozaki_i8.cuh)."""
    v = np.asarray(x, dtype=np.float64) / scale
    planes = []
    for _ in range(S):
        t = v * R
        d = np.rint(t)
        planes.append(d.astype(np.int64))
        v = t - d
    return planes


def split_product(x, a, S, R):
    """Y = X A^T through digit planes: groups g = k + l (0-based), pairs with k + l >= S dropped, Horner recombination."""
    sx = pow2_above(np.abs(x).max())
    sa = np.array([pow2_above(np.abs(r).max()) for r in a])
    dx = split_digits(x, sx, S, R)
    da = split_digits(a, sa[:, None], S, R)
    groups = [np.zeros((x.shape[0], a.shape[0]), dtype=np.int64) for _ in range(S)]
    for k in range(S):
        for l in range(S - k):
            groups[k + l] += dx[k] @ da[l].T
    acc = groups[S - 1].astype(np.float64)
    for g in range(S - 2, -1, -1):
        acc = acc / R + groups[g]
    return acc / (R * R) * sx * sa[None, :], dx, da, groups


@pytest.mark.parametrize("S,R,bound", [(7, 254, 127), (6, 254, 127), (5, 254, 127), (3, 254, 127), (6, 128, 64)])
def test_digit_ranges_and_int32_exactness(S, R, bound):
    rng = np.random.RandomState(S * 1000 + R)
    kmax = ((1 << 31) // ((R // 2) ** 2 * S)) // 64 * 64          # oz_kmax in csrc/host_session.cuh
    k = min(kmax, 4096)
    x = rng.randn(64, k) * 3.0
    x[3, 7] = 11.5
    a = rng.randn(9, k) * rng.uniform(1e-3, 10, size=(9, 1))
    y, dx, da, groups = split_product(x, a, S, R)
    assert max(int(np.abs(p).max()) for p in dx + da) <= bound    # int8 range, |d| <= R/2
    # worst case over a full-length contraction: (g+1) pairs per group, every product at most bound^2
    assert S * kmax * bound * bound < (1 << 31)
    assert max(int(np.abs(g).max()) for g in groups) < (1 << 31)


@pytest.mark.parametrize("S,R,tol", [(7, 254, 1e-14), (6, 254, 6e-13), (5, 254, 1.5e-10), (6, 128, 4e-11), (3, 254, 6e-6)])
def test_recombined_product_accuracy(S, R, tol):
    """Error relative to the natural scale max|x| * max_j|a_j| * sqrt(n): bounded by ~ R^-S (+ the dropped cross terms)."""
    rng = np.random.RandomState(11)
    n = 3000
    x = rng.randn(200, n)
    a = rng.randn(12, n) * rng.uniform(0.01, 5, size=(12, 1))
    a[:, ::5] *= 1e-4
    y, _, _, _ = split_product(x, a, S, R)
    ref = x @ a.T
    err = np.abs(y - ref).max(axis=0) / np.abs(ref).max(axis=0)
    assert err.max() < tol, err.max()
    # truncation is unbiased (rint): the mean signed error is far below the max error
    assert abs(np.mean((y - ref) / np.abs(ref).max(axis=0))) < 0.2 * tol


def test_radix_254_beats_radix_128_at_equal_cost():
    rng = np.random.RandomState(5)
    x, a = rng.randn(100, 2000), rng.randn(8, 2000)
    ref = x @ a.T
    e254 = np.abs(split_product(x, a, 6, 254)[0] - ref).max()
    e128 = np.abs(split_product(x, a, 6, 128)[0] - ref).max()
    assert e254 * 30 < e128


def mn_product(q, v, S, R, unit_diag):
    """out = Q V (Q m x m, V m x n) the way oz_mn_t in csrc/host_oz.cuh runs it: the contraction goes over V's ROWS, so V gets one
    power-of-two scale per COLUMN and Q one per row; with `unit_diag` the product runs on Q - I and V is added back exactly."""
    m = q.shape[0]
    qq = q - np.eye(m) if unit_diag else q
    sq = np.array([pow2_above(np.abs(r).max()) for r in qq])
    sv = np.array([pow2_above(np.abs(c).max()) for c in v.T])
    dq = split_digits(qq, sq[:, None], S, R)
    dv = split_digits(v, sv[None, :], S, R)
    groups = [np.zeros((m, v.shape[1]), dtype=np.int64) for _ in range(S)]
    for k in range(S):
        for l in range(S - k):
            groups[k + l] += dq[l] @ dv[k]
    acc = groups[S - 1].astype(np.float64)
    for g in range(S - 2, -1, -1):
        acc = acc / R + groups[g]
    out = acc / (R * R) * sq[:, None] * sv[None, :]
    return out + v if unit_diag else out


def test_column_scaled_product_and_unit_diagonal():
    """Qij = ry rinv (linearcorex.py:266) with ry = I + small correlations and rinv spanning orders of magnitude over the
    variables (rho -> 1 blows rinv up): per-column scales keep every variable at 48 bits, and taking the exact unit
    diagonal out of ry makes the digits resolve the 1e-2 off-diagonal entries instead of the 1."""
    rng = np.random.RandomState(3)
    m, n = 40, 900
    ry = rng.randn(m, m) * 1e-2
    ry = (ry + ry.T) / 2
    np.fill_diagonal(ry, 1.0)
    rinv = rng.randn(m, n) * 10.0 ** rng.uniform(-3, 2, size=(1, n))     # per-variable magnitude spread
    ref = ry @ rinv
    scale = np.abs(ref).max(axis=0)                                      # error measured per variable (column)
    with_diag = np.abs(mn_product(ry, rinv, 6, 254, False) - ref).max(axis=0) / scale
    no_diag = np.abs(mn_product(ry, rinv, 6, 254, True) - ref).max(axis=0) / scale
    assert with_diag.max() < 1e-12
    assert no_diag.max() < 5e-15
    assert no_diag.max() * 20 < with_diag.max()
    # one global scale for V instead of one per column would lose the small-magnitude variables altogether
    sv = pow2_above(np.abs(rinv).max())
    coarse = np.rint(rinv / sv * 254.0 ** 6) / 254.0 ** 6 * sv
    glob = np.abs(ry @ coarse - ref).max(axis=0) / scale
    assert glob.max() > 1e3 * no_diag.max()
