"""Wall time of full fits on the reference's own small configurations (README demo, big5, adni), where a
GPU iteration is launch- and synchronisation-bound.  GPU only.

    python tools/small_configs.py
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    import torch
    from conftest import load_golden
    from linearcorex_b200 import Corex
    for name in ("readme_demo_f64", "big5_l0_f64", "adni_l0_f64", "syn_4000x2000x20_f64"):
        z, kw, x = load_golden(name)
        for precision in ("fp64_split", "fp64"):
            Corex(precision=precision, **dict(kw, max_iter=2)).fit(x)  # warm-up
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            mdl = Corex(precision=precision, **kw).fit(x)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            its = len(mdl.history["TC"])
            trials = float(np.mean([r["trials"] for r in mdl.trace]))
            print("%-22s %-10s %5d iterations in %.3f s = %6.0f it/s (%.2f trials/iteration, %.3f ms/iteration)"
                  % (name, precision, its, dt, its / dt, trials, 1e3 * dt / its))


if __name__ == "__main__":
    main()
