"""pytest configuration: registers the `gpu` marker and exposes golden-vector helpers."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    """Load tests/golden/<name>.npz -> (dict of arrays, constructor kwargs, X)."""
    from corex_oracle import latent_factor_data
    z = dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))
    kwargs = {}
    for key in list(z):
        if key.startswith("kw_"):
            v = z[key]
            k = key[3:]
            if v.dtype.kind in "US":
                kwargs[k] = str(v)
            elif v.dtype.kind == "b":
                kwargs[k] = bool(v)
            elif v.dtype.kind == "f" and np.isnan(v):
                kwargs[k] = None
            elif v.dtype.kind in "iu":
                kwargs[k] = int(v)
            else:
                kwargs[k] = float(v)
    if "x_gen" in z:
        N, n, k, seed, snr, spread = z["x_gen"]
        x = latent_factor_data(int(N), int(n), int(k), seed=int(seed), snr=float(snr), snr_spread=float(spread))
    else:
        x = z.get("x")
    return z, kwargs, x


def golden_moments(z, prefix="m_"):
    return {k[len(prefix):]: v for k, v in z.items() if k.startswith(prefix)}


@pytest.fixture(scope="session")
def golden():
    return load_golden
