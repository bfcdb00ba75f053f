"""Per-iteration floor (launch + host-sync + Python overhead) measured on a problem too small to keep the GPU busy."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "oracle")
import corex_oracle as oc  # noqa: E402
from linearcorex_b200 import Corex  # noqa: E402

for prec in ("fp64", "fp64_split"):
    for (N, n, m) in ((256, 128, 8), (2000, 1000, 100)):
        x = oc.latent_factor_data(N, n, m, seed=0, snr=1.0, snr_spread=0.3)
        mdl = Corex(n_hidden=m, seed=0, tol=1e-12, max_iter=10 ** 9, precision=prec)
        sched = mdl._prepare(x)
        mdl._begin_stage(sched[0], rescale=False)
        for _ in range(20):
            mdl._iterate()
        torch.cuda.synchronize()
        l0 = mdl._sess.launches()
        t0 = time.perf_counter()
        K = 300
        for _ in range(K):
            mdl._iterate()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tr = np.mean([t["trials"] for t in mdl.trace[-K:]])
        print("%-10s %5dx%5dx%3d: %.1f us/iteration, %.1f launches/iteration, %.2f trials/iteration"
              % (prec, N, n, m, 1e6 * dt / K, (mdl._sess.launches() - l0) / K, tr), flush=True)
