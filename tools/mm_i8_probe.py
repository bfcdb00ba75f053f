"""Device time of the direction step and of a linear trial with the four m x m x n products on DMMA (LCX_MM_I8=0) and on
the int8 tcgen05 engine (LCX_MM_I8=1), plus the difference of the arrays they produce.
usage: python tools/mm_i8_probe.py N n m [precision [int8only]]      (JSON to stdout; `int8only` skips the DMMA pass, for ncu)"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from linearcorex_b200 import _lib  # noqa: E402
from linearcorex_b200.corex import _DeviceSession  # noqa: E402
from gpu_probe import timeit  # noqa: E402


def run(flag, N, nv, m, prec, xt, w):
    os.environ["LCX_MM_I8"] = flag
    sess = _DeviceSession(_lib.PRECISIONS[prec])
    lib = sess.lib
    sess.bind(xt.clone(), N, nv, m, None)
    sess.view(_lib.A_W).copy_(w)
    tc, muj, tang = C.c_double(), C.c_double(), C.c_double()
    _lib.check(lib.lcx_moments_ns(sess.h, 0.36, 0, C.byref(tc), C.byref(muj)))
    out = {"TC": tc.value, "max_uj": muj.value}
    l0 = sess.launches()
    _lib.check(lib.lcx_direction_ns(sess.h, 0.36, C.byref(tang)))
    out["direction_launches"] = sess.launches() - l0
    out["direction_ms"], _ = timeit(lambda: _lib.check(lib.lcx_direction_ns(sess.h, 0.36, C.byref(tang))))
    out["tangent"] = tang.value
    l0 = sess.launches()
    _lib.check(lib.lcx_trial_ns(sess.h, 0.36, 0.25, 0, C.byref(tc), C.byref(muj)))
    out["trial_launches"] = sess.launches() - l0
    out["trial_linear_ms"], _ = timeit(lambda: _lib.check(lib.lcx_trial_ns(sess.h, 0.36, 0.25, 0, C.byref(tc), C.byref(muj))))
    out["TC_trial"] = tc.value
    out["pair_ms"], _ = timeit(lambda: _lib.check(lib.lcx_sig(sess.h, sess.view(_lib.A_W).data_ptr(), 0.0,
                                                             sess.view(_lib.A_RDIR).data_ptr())))
    arrays = {k: sess.view(getattr(_lib, a), s).clone() for k, a, s in
              (("ry", "A_RY", 0), ("Qij", "A_QIJ", 0), ("grad", "A_GRAD", 0), ("update", "A_UPDATE", 0),
               ("ry_trial", "A_RY", 1), ("Qij_trial", "A_QIJ", 1))}
    # the two moment products against torch's fp64 matmul on the same operands
    W, rho, rinv = (sess.view(getattr(_lib, a)) for a in ("A_W", "A_RHO", "A_RHOINVRHO"))
    ry = W @ rho.t()
    ry.fill_diagonal_(1.0)
    out["ry_vs_torch"] = float(((arrays["ry"] - ry).abs().max() / ry.abs().max()).item())
    q = ry @ rinv
    out["Qij_vs_torch"] = float(((arrays["Qij"] - q).abs().max() / q.abs().max()).item())
    sess.close()
    return out, arrays


def main():
    N, nv, m = (int(v) for v in sys.argv[1:4])
    prec = sys.argv[4] if len(sys.argv) > 4 else "fp64_split"
    g = torch.Generator(device="cuda").manual_seed(0)
    k = max(2, m // 2)
    z = torch.randn(N, k, dtype=torch.float64, device="cuda", generator=g)
    parent = torch.arange(nv, device="cuda") % k
    xt = (z[:, parent] + torch.randn(N, nv, dtype=torch.float64, device="cuda", generator=g)) / np.sqrt(2.0)
    xt = (xt - xt.mean(0)) / xt.std(0, unbiased=False)
    ld = _lib.load().lcx_ld(nv)
    xp = torch.zeros(N, ld, dtype=torch.float64, device="cuda")
    xp[:, :nv] = xt
    w = torch.randn(m, nv, dtype=torch.float64, device="cuda", generator=g) / (10.0 * nv ** 0.5)
    res = {"shape": [N, nv, m], "precision": prec}
    if len(sys.argv) > 5 and sys.argv[5] == "int8only":
        res["int8"], _ = run("1", N, nv, m, prec, xp, w)
        print(json.dumps(res, indent=1))
        return
    a, arr_a = run("0", N, nv, m, prec, xp, w)
    b, arr_b = run("1", N, nv, m, prec, xp, w)
    res["dmma"], res["int8"] = a, b
    res["rel_diff"] = {k: float(((arr_a[k] - arr_b[k]).abs().max() / arr_a[k].abs().max()).item()) for k in arr_a}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
