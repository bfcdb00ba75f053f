"""Measured parity of the CUDA path against the golden vectors of the unmodified reference: max-norm relative error of W, the
TC trajectory, every moments key, transform, covariance and predict, per golden fit and precision mode.  Run on a B200:

    python tools/parity_report.py > profiles/r02_parity_report.txt

This is the evidence behind the tolerances in tests/ (every FP64-mode assertion is 1e-9)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
from conftest import load_golden, golden_moments  # noqa: E402
from linearcorex_b200 import Corex  # noqa: E402

CASES = ["readme_demo_f64", "big5_l0_f64", "big5_l1_f64", "syn_400x300x10_f64", "syn_60x400x8_f64",
         "syn_400x300x10_noanneal_f64", "outliers_missing_f64", "outliers_f64", "standard_missing_f64", "adni_l0_f64",
         "adni_l1_f64", "adni_l2_f64", "syn_4000x2000x20_f64", "syn_400x300x10_synergy_f64", "big5_syn_f64"]


def rel(a, b, floor=0.0):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if a.shape != b.shape:
        return float("nan")
    return float(np.abs(a - b).max() / max(np.abs(b).max(), floor, 1e-300))


def main():
    # a mode is a precision, optionally with the algorithm after a slash: fp64_split/gram = the Gram route
    modes = sys.argv[1:] or ["fp64", "fp64_split", "fp64_split/gram"]
    worst = {}
    for name in CASES:
        z, kw, x = load_golden(name)
        for mode in modes:
            prec, _, algo = mode.partition("/")
            if algo == "gram" and x is not None and np.shape(x)[0] < np.shape(x)[1]:
                continue  # (the Gram route is for N >= n)
            mdl = Corex(**dict(kw, precision=prec, algorithm=algo or "stream"))
            if name.startswith("readme_demo"):
                x = np.random.random((100, 50))
            mdl.fit(x)
            gm = golden_moments(z)
            tcs = abs(float(z["m_TC"]))
            errs = {"iters": "%d/%d" % (len(mdl.history["TC"]), len(z["history_TC"])), "ws": rel(mdl.ws, z["ws"])}
            n = min(len(mdl.history["TC"]), len(z["history_TC"]))
            errs["history"] = rel(mdl.history["TC"][:n], z["history_TC"][:n])
            for key, val in gm.items():
                if key in mdl.moments:
                    errs[key] = rel(mdl.moments[key], val, tcs if key in ("additivity", "TC_direct", "TC_no_overlap") else 0.0)
            errs["transform"] = rel(mdl.transform(x), z["transform"])
            if "covariance" in z:
                errs["covariance"] = rel(mdl.get_covariance(), z["covariance"])
            if "predict7" in z:
                errs["predict7"] = rel(mdl.predict(z["transform"][:7]), z["predict7"])
            errs["clusters_equal"] = bool((mdl.clusters() == z["clusters"]).all())
            top = sorted(((v, k) for k, v in errs.items() if isinstance(v, float)), reverse=True)[:4]
            print("%-30s %-16s iters %-10s clusters %s  worst: %s" % (
                name, mode, errs["iters"], errs["clusters_equal"], ", ".join("%s %.1e" % (k, v) for v, k in top)), flush=True)
            for k, v in errs.items():
                if isinstance(v, float):
                    worst[(mode, k)] = max(worst.get((mode, k), 0.0), v)
    print("\nworst case per key over all fits:")
    for mode in modes:
        print(mode, ", ".join("%s %.1e" % (k, v) for (mo, k), v in sorted(worst.items()) if mo == mode))


if __name__ == "__main__":
    main()
