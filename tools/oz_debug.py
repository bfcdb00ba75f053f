"""Debug probe for the split-integer tcgen05 contractions: compares lcx_sig (Y = X~ u^T, D = X~^T Y) in the split
modes against numpy float64 for a few shapes and prints max-norm relative errors."""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from linearcorex_b200 import _lib  # noqa: E402
from linearcorex_b200.corex import _DeviceSession  # noqa: E402


def run(N, n, m, precision, seed=0):
    rng = np.random.RandomState(seed)
    x = rng.randn(N, n)
    x[rng.randint(N), rng.randint(n)] = 7.3  # an outlier sets the global exponent
    u = rng.randn(m, n) * rng.uniform(0.01, 3.0, size=(m, 1)) / np.sqrt(n)
    u[:, ::3] *= 1e-3
    sess = _DeviceSession(precision)
    lib = sess.lib
    ld = lib.lcx_ld(n)
    xt = torch.zeros((N, ld), dtype=torch.float64, device="cuda")
    xt[:, :n] = torch.from_numpy(x)
    sess.bind(xt, N, n, m, None)
    ud = torch.zeros((m, ld), dtype=torch.float64, device="cuda")
    ud[:, :n] = torch.from_numpy(u)
    od = torch.zeros_like(ud)
    _lib.check(lib.lcx_sig(sess.h, ud.data_ptr(), 0.3, od.data_ptr()), "lcx_sig")
    torch.cuda.synchronize()
    y = sess.view(_lib.A_Y).cpu().numpy()
    y_ref = x @ u.T
    d_ref = (x.T @ y_ref).T
    sig_ref = (1 - 0.09) * d_ref / N + 0.09 * u
    got = od[:, :n].cpu().numpy()
    ey = np.abs(y - y_ref).max() / np.abs(y_ref).max()
    es = np.abs(got - sig_ref).max() / np.abs(sig_ref).max()
    # per-row (factor) relative error of Y, the scale the digits are relative to
    eyr = (np.abs(y - y_ref).max(0) / np.abs(y_ref).max(0)).max()
    print("N=%6d n=%6d m=%4d prec=%d  |dY|/|Y|=%.3e (per-factor %.3e)  |dsig|/|sig|=%.3e  nan=%s"
          % (N, n, m, precision, ey, eyr, es, bool(np.isnan(got).any() or np.isnan(y).any())), flush=True)
    sess.close()
    return ey, es


if __name__ == "__main__":
    shapes = [(128, 64, 64), (256, 128, 64), (300, 200, 10), (1000, 333, 100), (4000, 2000, 20), (513, 1000, 130),
              (20000, 5000, 100), (70000, 300, 5)]
    for prec in (2, 1):
        for (N, n, m) in shapes:
            run(N, n, m, prec)
