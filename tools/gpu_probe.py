"""One-off GPU probe: FP64 / TF32 library GEMM peaks (roofline denominators that MEASURED_PEAKS.json lacks)
and the device time of the two X contractions at a given shape.  Writes JSON to stdout."""
import ctypes as C
import json
import sys

import torch

sys.path.insert(0, ".")
from linearcorex_b200 import _lib  # noqa: E402
from linearcorex_b200.corex import _DeviceSession  # noqa: E402


def timeit(fn, warm=2, iters=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best, tot = 1e30, 0.0
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = min(best, ms)
        tot += ms
    return best, tot / iters


def main():
    out = {"gpu": torch.cuda.get_device_name(0)}
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    best, mean = timeit(lambda: torch.matmul(a, b))
    out["cublas_dgemm_tflops_best"] = 2 * n ** 3 / best / 1e9
    out["cublas_dgemm_tflops_mean"] = 2 * n ** 3 / mean / 1e9
    # the skinny shapes of the hot path through cuBLAS, for reference
    N, nv, m = (int(v) for v in (sys.argv[1:4] if len(sys.argv) >= 4 else (100000, 10000, 100)))
    x = torch.randn(N, nv, dtype=torch.float64, device="cuda")
    w = torch.randn(m, nv, dtype=torch.float64, device="cuda")
    best, mean = timeit(lambda: torch.matmul(x, w.t()))
    out["cublas_k1_ms"] = best
    out["cublas_k1_tflops"] = 2.0 * N * nv * m / best / 1e9
    y = torch.matmul(x, w.t())
    best, mean = timeit(lambda: torch.matmul(x.t(), y))
    out["cublas_k2_ms"] = best
    out["cublas_k2_tflops"] = 2.0 * N * nv * m / best / 1e9
    del a, b
    a32 = torch.randn(n, n, dtype=torch.float32, device="cuda")
    b32 = torch.randn(n, n, dtype=torch.float32, device="cuda")
    torch.backends.cuda.matmul.allow_tf32 = True
    best, mean = timeit(lambda: torch.matmul(a32, b32), iters=10)
    out["cublas_tf32_tflops_best"] = 2 * n ** 3 / best / 1e9
    torch.backends.cuda.matmul.allow_tf32 = False
    best, mean = timeit(lambda: torch.matmul(a32, b32))
    out["cublas_sgemm_tflops_best"] = 2 * n ** 3 / best / 1e9
    del a32, b32

    prec = int(sys.argv[4]) if len(sys.argv) >= 5 else _lib.PRECISION_FP64
    out["precision"] = prec
    sess = _DeviceSession(prec)
    lib = sess.lib
    ld = lib.lcx_ld(nv)
    assert ld == nv or True
    xt = torch.zeros(N, ld, dtype=torch.float64, device="cuda")
    xt[:, :nv] = x
    del x
    sess.bind(xt, N, nv, m, None)
    wv = sess.view(_lib.A_W)
    wv.copy_(w / (10 * nv ** 0.5))
    tc, muj, tang = C.c_double(), C.c_double(), C.c_double()
    l0 = sess.launches()
    best, mean = timeit(lambda: _lib.check(lib.lcx_sig(sess.h, sess.view(_lib.A_W).data_ptr(), 0.0,
                                                      sess.view(_lib.A_GRAD).data_ptr())), iters=5)
    out["lcx_pair_ms_best"] = best
    out["lcx_pair_ms_mean"] = mean
    out["lcx_pair_tflops"] = 4.0 * N * nv * m / best / 1e9
    out["lcx_pair_launches"] = (sess.launches() - l0) / 7
    # K1 alone
    ldy = lib.lcx_ldy(m)
    yb = torch.empty(N, ldy, dtype=torch.float64, device="cuda")
    best, mean = timeit(lambda: _lib.check(lib.lcx_project(sess.h, xt.data_ptr(), N, nv, ld, wv.data_ptr(), ld, m,
                                                          yb.data_ptr(), ldy, None, None, 0)))
    out["lcx_k1_ms"] = best
    out["lcx_k1_tflops"] = 2.0 * N * nv * m / best / 1e9
    out["lcx_k2_ms_est"] = out["lcx_pair_ms_best"] - best
    # full iteration pieces
    _lib.check(lib.lcx_moments_ns(sess.h, 0.0, 0, C.byref(tc), C.byref(muj)))
    best, mean = timeit(lambda: _lib.check(lib.lcx_direction_ns(sess.h, 0.0, C.byref(tang))))
    out["lcx_direction_ms"] = best
    best, mean = timeit(lambda: _lib.check(lib.lcx_trial_ns(sess.h, 0.0, 1e-3, 0, C.byref(tc), C.byref(muj))))
    out["lcx_trial_linear_ms"] = best
    best, mean = timeit(lambda: _lib.check(lib.lcx_trial_ns(sess.h, 0.0, 1e-3, 1, C.byref(tc), C.byref(muj))))
    out["lcx_trial_exact_ms"] = best
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
