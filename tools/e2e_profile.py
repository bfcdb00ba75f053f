"""Where the end-to-end `Corex.fit(pinned host X)` time goes at config 3, and how many iterations a fit with the
reference's default stopping rule (tol=1e-5) takes there.  GPU only; run under gpurun.

    python tools/e2e_profile.py [--rows 100000] [--vars 10000] [--factors 100] [--per-stage 6] [--converge 400]
"""
import argparse
import cProfile
import io
import os
import pstats
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=100000)
    ap.add_argument("--vars", type=int, default=10000)
    ap.add_argument("--factors", type=int, default=100)
    ap.add_argument("--per-stage", type=int, default=6)
    ap.add_argument("--converge", type=int, default=400, help="max_iter of the tol=1e-5 fit (0 = skip)")
    ap.add_argument("--precision", default="fp64_split")
    args = ap.parse_args()
    import torch
    import bench
    from linearcorex_b200 import Corex
    x = bench.make_rows(args.rows, args.vars, args.factors, 0, args.rows)
    xp = torch.from_numpy(x).pin_memory()

    def one(max_iter, tol=1e-12):
        mdl = Corex(n_hidden=args.factors, seed=0, tol=tol, max_iter=max_iter, precision=args.precision)
        t0 = time.perf_counter()
        mdl.fit(xp)
        torch.cuda.synchronize()
        return mdl, time.perf_counter() - t0

    one(1)  # warm-up: module load, allocator pools
    mdl, dt = one(args.per_stage)
    its = len(mdl.history["TC"])
    print("fit(max_iter=%d): %d iterations in %.4f s = %.1f it/s; phases %s" % (args.per_stage, its, dt, its / dt, mdl.timings))
    del mdl
    pr = cProfile.Profile()
    pr.enable()
    mdl, dt = one(args.per_stage)
    pr.disable()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28)
    print(s.getvalue())
    del mdl
    if args.converge:
        mdl, dt = one(args.converge, tol=1e-5)
        its = len(mdl.history["TC"])
        stages = {}
        for rec in mdl.trace:
            stages.setdefault(rec["eps"], []).append(rec["trials"])
        print("fit(tol=1e-5, max_iter=%d): %d iterations in %.3f s = %.1f it/s; TC=%.6f; phases %s"
              % (args.converge, its, dt, its / dt, mdl.tc, mdl.timings))
        for eps, tr in stages.items():
            print("  eps=%.6f: %d iterations, %.2f trials/iteration" % (eps, len(tr), float(np.mean(tr))))


if __name__ == "__main__":
    main()
