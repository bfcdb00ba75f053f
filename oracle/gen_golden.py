"""Generate tests/golden/*.npz by running the UNMODIFIED reference (TEST INFRASTRUCTURE ONLY).

Run in the build container, where `/root/reference` is mounted:

    python oracle/gen_golden.py            # writes tests/golden/*.npz

The reference cannot travel to the GPU box, so the vectors are committed.  Nothing under
`tests/`, `bench.py` or the package reads `/root/reference` at run time.

Modes:
  f64     the reference with its module-global `np` replaced by a shim whose `float32` is
          `float64` -- turns the casts at linearcorex.py:108 and :116 into float64 casts.
          This is "the reference's numpy float64 path", the parity target of the device FP64 mode.
  native  the reference exactly as shipped (float32 casts).
The CLI path (`missing_values=-1e6`) is simply `native` + that kwarg.

Each golden holds: the input X (or the generator arguments), constructor kwargs, the final `ws`,
every `moments` key, `theta`, `history['TC']`, per-iteration trial counts, `transform(X)`,
`get_covariance()`, `clusters()`; "step" goldens hold one `_calculate_moments_*`/`_sig`/`_update_*`
call's inputs and outputs.
"""
import os
import sys

import numpy as np

REF_ROOT = os.environ.get("LCX_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
sys.path.insert(0, HERE)
sys.path.insert(0, REF_ROOT)

import linearcorex.linearcorex as ref  # noqa: E402  (the unmodified reference)
from corex_oracle import latent_factor_data  # noqa: E402


class _F64Shim(object):
    """numpy stand-in whose float32 is float64; everything else delegates to numpy."""
    float32 = np.float64

    def __getattr__(self, name):
        return getattr(np, name)


class _Traced(ref.Corex):
    """Counts quick-moment trials per `_update_ns` call without editing the reference."""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.trials, self.quick_fails, self._in_update = [], [], False

    def _calculate_moments_ns(self, x, ws, quick=False):
        out = super()._calculate_moments_ns(x, ws, quick=quick)
        if self._in_update and quick:
            self.trials[-1] += 1
            if out is False:
                self.quick_fails[-1] += 1
        return out

    def _update_ns(self, x):
        self.trials.append(0)
        self.quick_fails.append(0)
        self._in_update = True
        try:
            return super()._update_ns(x)
        finally:
            self._in_update = False


def _with_mode(mode, fn):
    if mode == "f64":
        ref.np = _F64Shim()
        try:
            return fn()
        finally:
            ref.np = np
    return fn()


def _pack_moments(prefix, moments, out):
    for key, val in moments.items():
        out[prefix + key] = np.asarray(val)


def run_fit(name, x, mode, kwargs, pre_ctor_draw=None, cov=True, x_gen=None):
    """Fit the reference on x and dump everything observable."""
    def go():
        mdl = _Traced(**kwargs)
        xx = x
        if pre_ctor_draw is not None:  # README demo draws X *after* the constructor seeded the RNG
            xx = pre_ctor_draw()
        mdl.fit(xx)
        return mdl, xx
    mdl, xx = _with_mode(mode, go)
    out = {"mode": mode, "kwargs_keys": np.array(sorted(kwargs.keys()))}
    for k, v in kwargs.items():
        out["kw_" + k] = np.asarray(np.nan if v is None else v)
    if x_gen is not None:
        out["x_gen"] = np.asarray(x_gen, dtype=np.float64)
    else:
        out["x"] = np.asarray(xx)
    out["ws"] = mdl.ws
    out["theta_mean"], out["theta_std"] = np.asarray(mdl.theta[0]), np.asarray(mdl.theta[1])
    out["n_obs"] = np.asarray(mdl.n_obs)
    out["history_TC"] = np.asarray(mdl.history.get("TC", []), dtype=np.float64)
    out["trials"] = np.asarray(mdl.trials)
    out["quick_fails"] = np.asarray(mdl.quick_fails)
    out["clusters"] = mdl.clusters()
    out["eps_final"] = np.asarray(mdl.eps)
    _pack_moments("m_", mdl.moments, out)
    out["mis"] = mdl.mis

    def tr():
        return mdl.transform(xx)
    out["transform"] = _with_mode(mode, tr)
    if cov:
        out["covariance"] = mdl.get_covariance()
    if "X_i Z_j" in mdl.moments:
        yy = out["transform"][:7]
        out["predict7"] = mdl.predict(yy)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("%-28s mode=%-6s it=%5d TC=%.12g  ws=%s" % (name, mode, len(out["history_TC"]),
                                                       float(mdl.tc), mdl.ws.dtype))
    return mdl, out


def run_step(name, x, mode, kwargs, n_warm=3):
    """Single-call goldens: moments (quick/full), _sig, one _update step, from a state a few
    iterations into a fit (so rho etc. are non-trivial)."""
    def go():
        mdl = _Traced(**dict(kwargs, max_iter=n_warm, anneal=False))
        mdl.fit(x)  # leaves eps = 0, sorted ws, full moments
        xt = mdl.preprocess(np.asarray(x, dtype=ref.np.float32))
        out = {"mode": mode, "x": np.asarray(x), "xt": xt, "w": mdl.ws.copy()}
        for k, v in kwargs.items():
            out["kw_" + k] = np.asarray(np.nan if v is None else v)
        for eps in (0.0, 0.36):
            mdl.eps = eps
            tag = "e%02d_" % int(round(eps * 100))
            mq = mdl._calculate_moments(xt, mdl.ws, quick=True)
            mf = mdl._calculate_moments(xt, mdl.ws, quick=False)
            _pack_moments(tag + "q_", mq, out)
            _pack_moments(tag + "f_", mf, out)
            u = np.cos(np.arange(mdl.ws.size, dtype=np.float64)).reshape(mdl.ws.shape).astype(mdl.ws.dtype)
            out[tag + "sig_u"] = u
            out[tag + "sig"] = mdl._sig(xt, u)
            out[tag + "norm"] = mdl._norm(xt, mdl.ws)
            mdl.moments = mf
            if mdl.discourage_overlap:
                w2, m2 = mdl._update_ns(xt)
                out[tag + "trials"] = np.asarray(mdl.trials[-1])
            else:
                w2, m2 = mdl._update_syn(xt, eta=0.1)
            out[tag + "w_next"] = w2
            _pack_moments(tag + "n_", m2, out)
        return out
    out = _with_mode(mode, go)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("%-28s mode=%-6s (step golden)" % (name, mode))


def load_csv(path, skip_first_col):
    with open(path, "r", newline=None) as fh:  # universal newlines: big5 uses CR-only endings
        rows = [r for r in fh.read().splitlines() if r.strip()]
    body = [r.split(",")[(1 if skip_first_col else 0):] for r in rows[1:]]
    return np.array(body, dtype=float)


def main():
    os.makedirs(OUT, exist_ok=True)
    data = os.path.join(REF_ROOT, "tests", "data")
    big5 = load_csv(os.path.join(data, "test_big5.csv"), skip_first_col=False)
    tiny = load_csv(os.path.join(data, "test_data.csv"), skip_first_col=True)
    adni = load_csv(os.path.join(data, "adni_blood.csv"), skip_first_col=True)
    assert big5.shape == (2000, 50) and tiny.shape == (8, 5) and adni.shape == (566, 200)

    # config 1: README demo (README.md:49-51) -- X drawn after the constructor seeded the RNG
    for mode in ("f64", "native"):
        run_fit("readme_demo_" + mode, None, mode, dict(n_hidden=5, seed=0),
                pre_ctor_draw=lambda: np.random.random((100, 50)))

    # config 2: big5, layers 5,1 (README.md:38)
    for mode in ("f64", "native"):
        mdl, out = run_fit("big5_l0_" + mode, big5, mode, dict(n_hidden=5, seed=0))
        run_fit("big5_l1_" + mode, out["transform"], mode, dict(n_hidden=1, seed=0))
    mdl, out = run_fit("big5_l0_cli", big5, "native", dict(n_hidden=5, seed=0, missing_values=-1e6))
    run_fit("big5_syn_f64", big5, "f64", dict(n_hidden=5, seed=0, discourage_overlap=False, max_iter=300))

    # 8x5 toy with the CLI's missing marker
    run_fit("test_data_cli", tiny, "native", dict(n_hidden=2, seed=0, missing_values=-1e6))
    run_fit("test_data_f64", tiny, "f64", dict(n_hidden=2, seed=0, missing_values=-1e6))

    # adni: 3.1 % missing entries, layers 30,5,1 (README.md:39); f64 + CLI path
    x_prev = adni
    for depth, width in enumerate((30, 5, 1)):
        kw = dict(n_hidden=width, seed=0)
        if depth == 0:
            kw["missing_values"] = -1e6
        mdl, out = run_fit("adni_l%d_f64" % depth, x_prev, "f64", kw, cov=(depth == 0))
        x_prev = out["transform"]
    run_fit("adni_l0_cli", adni, "native", dict(n_hidden=30, seed=0, missing_values=-1e6), cov=False)

    # synthetic modular data, well-separated TCs (snr_spread) -- X regenerated from x_gen
    for tag, (N, n, k, spread) in {"syn_400x300x10": (400, 300, 10, 0.35),
                                  "syn_60x400x8": (60, 400, 8, 0.5),
                                  "syn_4000x2000x20": (4000, 2000, 20, 0.2)}.items():
        x = latent_factor_data(N, n, k, seed=0, snr=1.0, snr_spread=spread)
        big = N * n > 1e6
        run_fit(tag + "_f64", x, "f64", dict(n_hidden=k, seed=0), cov=not big,
                x_gen=(N, n, k, 0, 1.0, spread))
    x = latent_factor_data(400, 300, 10, seed=0, snr=1.0, snr_spread=0.35)
    run_fit("syn_400x300x10_native", x, "native", dict(n_hidden=10, seed=0), x_gen=(400, 300, 10, 0, 1.0, 0.35))
    run_fit("syn_400x300x10_noanneal_f64", x, "f64", dict(n_hidden=10, seed=0, anneal=False),
            x_gen=(400, 300, 10, 0, 1.0, 0.35))
    run_fit("syn_400x300x10_synergy_f64", x, "f64",
            dict(n_hidden=10, seed=0, discourage_overlap=False, max_iter=200), x_gen=(400, 300, 10, 0, 1.0, 0.35))

    # outliers + missing: heavy-tailed columns, NaN marker
    rng = np.random.RandomState(7)
    xo = latent_factor_data(300, 120, 6, seed=3, snr=2.0, snr_spread=0.4).astype(np.float64)
    xo[:, ::7] = np.sign(xo[:, ::7]) * np.abs(xo[:, ::7]) ** 3  # long tails
    holes = rng.rand(*xo.shape) < 0.04
    xo_marker = np.where(holes, -1e6, xo)
    run_fit("outliers_missing_f64", xo_marker, "f64",
            dict(n_hidden=6, seed=0, gaussianize="outliers", missing_values=-1e6))
    run_fit("outliers_f64", xo, "f64", dict(n_hidden=6, seed=0, gaussianize="outliers"))
    run_fit("standard_missing_f64", xo_marker, "f64", dict(n_hidden=6, seed=0, missing_values=-1e6))

    # single-call goldens
    run_step("step_ns_400x300x10_f64", x, "f64", dict(n_hidden=10, seed=0))
    run_step("step_syn_400x300x10_f64", x, "f64", dict(n_hidden=10, seed=0, discourage_overlap=False))
    run_step("step_ns_big5_f64", big5, "f64", dict(n_hidden=5, seed=0))
    run_step("step_ns_60x400x8_f64", latent_factor_data(60, 400, 8, seed=0, snr=1.0, snr_spread=0.5), "f64",
             dict(n_hidden=8, seed=0))


if __name__ == "__main__":
    main()
