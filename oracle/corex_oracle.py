"""CPU oracle for the Linear CorEx fit loop -- TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of the algorithm implemented by the reference
(gregversteeg/LinearCorex, `linearcorex/linearcorex.py`).  It exists so that the
CUDA path in `linearcorex_b200/` can be checked on a box where `/root/reference`
is absent.  It is the *checker*, never the thing shipped or measured: only
`tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py` may import it.  The product package must never import it.

Parity status: PINNED.  `oracle/gen_golden.py` imports the unmodified reference
in the build container and writes `tests/golden/*.npz`; `tests/test_oracle_golden.py`
asserts that every function below reproduces those vectors (the reference itself
ships no tests or golden vectors -- SURVEY.md section 4).

Structure (all citations are `linearcorex/linearcorex.py:<line>` of the reference):

  standardize / impute_missing / squash_tails ........ :397-429, :483-510
  project_sumsq / norm_y ............................. :215-228, :247-249
  sigma_times ........................................ :196-213
  moments_ns ......................................... :236-288
  step_ns (direction + backtracking) ................. :290-334
  moments_syn / step_syn ............................. :336-384
  OracleCorex.fit / transform / get_covariance ....... :107-164, :386-395, :443-455
  fit_layers (the CLI's layer-stacking loop) ......... vis_corex.py:529-545

dtype policy: the reference casts the input to float32 (:108) and draws W in
float32 (:116); several of its paths silently promote to float64 (SURVEY.md
section 0.1).  `work_dtype` selects the cast applied at those two places:
`np.float32` reproduces the reference as shipped ("O-native" / "O-cli"),
`np.float64` reproduces the all-float64 run ("O-f64", the parity target of the
FP64 device mode).  Everything downstream follows numpy promotion exactly as the
same expressions do in the reference.
"""
from __future__ import annotations

import numpy as np

ANNEAL_BASE = 0.6
ANNEAL_STAGES = 6


# --------------------------------------------------------------------------------------
# preprocessing (:397-429, :483-510)
# --------------------------------------------------------------------------------------
def impute_missing(x, marker):
    """Column-mean imputation of entries equal to `marker` (or NaN).  (:497-510)

    Returns (imputed copy, n_obs per column as int array)."""
    x = np.array(x, copy=True)
    if not np.isnan(marker):
        x = np.where(x == marker, np.nan, x)
    cols, counts = [], []
    for col in x.T:
        col = np.array(col, copy=True)
        ok = np.isfinite(col)
        col[np.isnan(col)] = np.mean(col[ok])
        cols.append(col)
        counts.append(int(ok.sum()))
    return np.array(cols).T, np.array(counts)


def squash_tails(z, t=4):
    """g(): identity inside [-t, t], tanh-compressed outside.  (:483-487)"""
    core = np.clip(z, -t, t)
    return core + np.tanh(z - core)


def unsquash_tails(z, t=4):
    """g_inv().  (:490-494)"""
    core = np.clip(z, -t, t)
    return core + np.arctanh(np.clip(z - core, -1 + 1e-10, 1 - 1e-10))


def standardize(x, gaussianize, missing_values, theta=None):
    """preprocess() (:397-429).  theta=None means `fit=True`.

    Returns (x_tilde, theta, n_obs)."""
    if missing_values is not None:
        x, n_obs = impute_missing(x, missing_values)
    else:
        n_obs = len(x)
    if gaussianize == 'standard':
        if theta is None:
            mu = np.mean(x, axis=0)
            sd = np.sqrt(np.sum((x - mu) ** 2, axis=0) / n_obs).clip(1e-10)
            theta = (mu, sd)
        x = (x - theta[0]) / theta[1]
    elif gaussianize == 'outliers':
        if theta is None:
            mu = np.mean(x, axis=0)
            sd = np.std(x, axis=0, ddof=0).clip(1e-10)
            theta = (mu, sd)
        x = squash_tails((x - theta[0]) / theta[1])
    elif gaussianize == 'none':
        pass
    else:
        raise ValueError("oracle supports gaussianize in {'standard','outliers','none'}; "
                         "'empirical' is out of scope (SURVEY.md section 2 #8)")
    return x, theta, n_obs


# --------------------------------------------------------------------------------------
# the two X contractions
# --------------------------------------------------------------------------------------
def project_sumsq(xt, a):
    """Y = X~ A^T and s_j = sum_l Y_lj^2.  (:247-248, :226-227)"""
    y = xt.dot(a.T)
    return y, np.einsum('lj,lj->j', y, y)


def sigma_times(xt, u, eps):
    """_sig(): (1-eps^2) (X~^T (X~ u^T))^T / N + eps^2 u, an m x n array.  (:196-213)"""
    n_samples = xt.shape[0]
    y = xt.dot(u.T)
    d = xt.T.dot(y)
    return (1 - eps ** 2) * d.T / n_samples + eps ** 2 * u


def norm_y(xt, w, eps):
    """_norm(): sqrt(uj).  (:215-228)"""
    _, s = project_sumsq(xt, w)
    return np.sqrt((1 - eps ** 2) * s / xt.shape[0] + eps ** 2 * np.sum(w ** 2, axis=1))


# --------------------------------------------------------------------------------------
# non-synergistic moments and update
# --------------------------------------------------------------------------------------
def moments_ns(xt, w, eps, quick=False, yscale=1.):
    """_calculate_moments_ns (:236-288).  Returns None where the reference returns False."""
    n_samples = xt.shape[0]
    y, s = project_sumsq(xt, w)
    m = {}
    m["uj"] = (1 - eps ** 2) * s / n_samples + eps ** 2 * np.sum(w ** 2, axis=1)
    if quick and np.max(m["uj"]) >= 1.:
        return None
    d = xt.T.dot(y)
    rho = (1 - eps ** 2) * d.T / n_samples + eps ** 2 * w
    ry = w.dot(rho.T)
    m["Y_j^2"] = yscale ** 2 / (1. - m["uj"])
    np.fill_diagonal(ry, 1)
    inv = 1. / (1. - rho ** 2)
    rinv = rho * inv
    qij = np.dot(ry, rinv)
    si = np.sum(rho * rinv, axis=0)
    qs = np.einsum('ki,ki->i', rinv, qij - si * rho)
    m["rho"], m["ry"], m["invrho"], m["rhoinvrho"] = rho, ry, inv, rinv
    m["Qij"], m["Si"], m["Qi-Si^2"] = qij, si, qs
    m["TC"] = np.sum(np.log(1 + si)) - 0.5 * np.sum(np.log(1 + qs)) + 0.5 * np.sum(np.log(1 - m["uj"]))
    if not quick:
        m["MI"] = - 0.5 * np.log1p(-rho ** 2)
        m["X_i Y_j"] = rho.T * np.sqrt(m["Y_j^2"])
        m["X_i Z_j"] = np.linalg.solve(ry, rho).T
        m["X_i^2 | Y"] = (1. - np.einsum('ij,ji->i', m["X_i Z_j"], rho)).clip(1e-6)
        m['I(Y_j ; X)'] = 0.5 * np.log(m["Y_j^2"]) - 0.5 * np.log(yscale ** 2)
        m['I(X_i ; Y)'] = - 0.5 * np.log(m["X_i^2 | Y"])
        m["TCs"] = m["MI"].sum(axis=1) - m['I(Y_j ; X)']
        m["TC_no_overlap"] = m["MI"].max(axis=0).sum() - m['I(Y_j ; X)'].sum()
        m["TC_direct"] = m['I(X_i ; Y)'].sum() - m['I(Y_j ; X)']
        m["additivity"] = (m["MI"].sum(axis=0) - m['I(X_i ; Y)']).sum()
    return m


def direction_ns(xt, w, m, eps):
    """Search direction of _update_ns (:292-305): returns (grad, sig_grad, update, tangent)."""
    rj = 1. - m["uj"][:, np.newaxis]
    h = np.dot(m["rhoinvrho"] / (1 + m["Qi-Si^2"]), m["rhoinvrho"].T)
    np.fill_diagonal(h, 0)
    grad = w / rj
    grad -= 2 * m["invrho"] * m["rhoinvrho"] / (1 + m["Si"])
    grad += m["invrho"] ** 2 * ((1 + m["rho"] ** 2) * m["Qij"] - 2 * m["rho"] * m["Si"]) / (1 + m["Qi-Si^2"])
    grad += np.dot(h, w)
    sig_grad = sigma_times(xt, grad, eps)
    bj = np.sum(m["rho"] * grad, axis=1, keepdims=True)
    update = - rj * (grad - 2. * w / (2 - rj) * bj)
    tangent = np.einsum('ji,ji', sig_grad, update)
    return grad, sig_grad, update, tangent


def step_ns(xt, w, m, eps, tol, trace=None):
    """One _update_ns (:290-334).  Returns (w_new, moments_new); moments_new may be None
    ("step too small" quirk, :316-319).  `trace`, if a dict, receives eta / trials / quick-fails."""
    _, _, update, tangent = direction_ns(xt, w, m, eps)
    if trace is not None:
        trace.update(tangent=float(tangent), eta=0.0, trials=0, quick_fails=0)
    if tangent >= 0:
        return w, m
    eta = 1.
    w_try, m_try = None, None
    while True:
        if eta < min(tol, 1e-10):
            break
        w_try = w + eta * update
        m_try = moments_ns(xt, w_try, eps, quick=True)
        if trace is not None:
            trace["trials"] += 1
        if m_try is None:
            if trace is not None:
                trace["quick_fails"] += 1
            eta *= 0.5
            continue
        if not (-m_try['TC'] <= -m['TC'] + 0.1 * eta * tangent):
            eta *= 0.5
            continue
        break
    if trace is not None:
        trace["eta"] = eta
    return w_try, m_try


# --------------------------------------------------------------------------------------
# synergistic variant
# --------------------------------------------------------------------------------------
def moments_syn(xt, w, yscale=1.):
    """_calculate_moments_syn (:336-373); no eps terms, `quick` ignored by the reference."""
    n_samples = xt.shape[0]
    mdim = w.shape[0]
    m = {}
    y = xt.dot(w.T)
    xy = xt.T.dot(y) / n_samples
    cy = w.dot(xy) + yscale ** 2 * np.eye(mdim)
    yj2 = np.diag(cy).copy()
    m["X_i Y_j"], m["cy"], m["Y_j^2"] = xy, cy, yj2
    m["ry"] = cy / (np.sqrt(yj2) * np.sqrt(yj2[:, np.newaxis]))
    rho = (xy / np.sqrt(yj2)).T
    inv = 1. / (1. - rho ** 2)
    rinv = rho * inv
    m["rho"], m["invrho"], m["rhoinvrho"] = rho, inv, rinv
    m["Qij"] = np.dot(m['ry'], rinv)
    m["Qi"] = np.einsum('ki,ki->i', rinv, m["Qij"])
    m["Si"] = np.sum(rho * rinv, axis=0)
    m["MI"] = - 0.5 * np.log1p(-rho ** 2)
    m["X_i Z_j"] = np.linalg.solve(cy, xy.T).T
    m["X_i^2 | Y"] = (1. - np.einsum('ij,ij->i', m["X_i Z_j"], xy)).clip(1e-6)
    mi_y = 0.5 * np.log(yj2) - 0.5 * np.log(yscale ** 2)
    mi_x = - 0.5 * np.log(m["X_i^2 | Y"])
    m["TCs"] = m["MI"].sum(axis=1) - mi_y
    m["additivity"] = (m["MI"].sum(axis=0) - mi_x).sum()
    m["TC"] = np.sum(mi_x) - np.sum(mi_y)
    return m


def step_syn(xt, w, m, eta=0.1):
    """_update_syn (:375-384)."""
    h = (1. / m["X_i^2 | Y"] * m["X_i Z_j"].T).dot(m["X_i Z_j"])
    np.fill_diagonal(h, 0)
    r = m["X_i Z_j"].T / m["X_i^2 | Y"]
    w_new = (1. - eta) * w + eta * (r - np.dot(h, w))
    return w_new, moments_syn(xt, w_new)


# --------------------------------------------------------------------------------------
# the model driver
# --------------------------------------------------------------------------------------
class OracleCorex(object):
    """Restatement of `Corex` (:22-455) with the same constructor kwargs (:72-74)."""

    def __init__(self, n_hidden=10, max_iter=10000, tol=1e-5, anneal=True, missing_values=None,
                 discourage_overlap=True, gaussianize='standard', verbose=False, seed=None,
                 work_dtype=np.float32):
        self.m = n_hidden
        self.max_iter, self.tol, self.anneal = max_iter, tol, anneal
        self.missing_values = missing_values
        self.discourage_overlap = discourage_overlap
        self.gaussianize = gaussianize
        self.verbose = verbose
        self.work_dtype = work_dtype
        self.eps = 0
        self.yscale = 1.
        np.random.seed(seed)  # :89 -- global legacy RNG, seeded at construction
        self.n_samples, self.nv = 0, 0
        self.ws = np.zeros((0, 0))
        self.moments = {}
        self.theta = None
        self.history = {}
        self.trace = []  # per-iteration dicts: eps, eta, trials, quick_fails, tangent, TC

    # -- helpers ---------------------------------------------------------------------
    def _moments(self, xt, w, quick=False):
        if self.discourage_overlap:
            return moments_ns(xt, w, self.eps, quick=quick, yscale=self.yscale)
        return moments_syn(xt, w, yscale=self.yscale)

    def preprocess(self, x, fit=False):
        xt, theta, n_obs = standardize(x, self.gaussianize, self.missing_values,
                                       theta=None if fit else self.theta)
        if fit:
            self.theta = theta
        self.n_obs = n_obs
        return xt

    @property
    def tc(self):
        return self.moments["TC"]

    @property
    def tcs(self):
        return self.moments["TCs"]

    @property
    def mis(self):
        return - 0.5 * np.log1p(-self.moments["rho"] ** 2)

    def clusters(self):
        return np.argmax(np.abs(self.ws), axis=0)

    # -- fit (:107-164) ----------------------------------------------------------------
    def fit(self, x):
        xt = self.preprocess(np.asarray(x, dtype=self.work_dtype), fit=True)
        self.n_samples, self.nv = xt.shape
        schedule = [0.]
        if self.ws.size == 0:
            if self.discourage_overlap:
                self.ws = np.random.randn(self.m, self.nv).astype(self.work_dtype)
                self.ws /= (10. * norm_y(xt, self.ws, self.eps))[:, np.newaxis]
                if self.anneal:
                    schedule = [ANNEAL_BASE ** k for k in range(1, ANNEAL_STAGES + 1)] + [0]
            else:
                self.ws = np.random.randn(self.m, self.nv) * self.yscale ** 2 / np.sqrt(self.nv)
        self.moments = self._moments(xt, self.ws, quick=True)

        for stage, eps in enumerate(schedule):
            eps_prev, self.eps = self.eps, eps
            if stage > 0:  # :129-133
                wmag = np.sum(self.ws ** 2, axis=1, keepdims=True)
                delta = (eps ** 2 - eps_prev ** 2) / (1. - eps ** 2) * wmag / self.moments['uj'].reshape((-1, 1))
                a = np.sqrt((1. - eps_prev ** 2) / ((1. - eps ** 2) * (1. + delta)))
                self.ws *= 0.001 * np.floor(1000. * a)
            self.moments = self._moments(xt, self.ws)
            for _ in range(self.max_iter):
                last_tc = self.tc
                rec = {"eps": eps}
                if self.discourage_overlap:
                    self.ws, self.moments = step_ns(xt, self.ws, self.moments, eps, self.tol, trace=rec)
                else:
                    self.ws, self.moments = step_syn(xt, self.ws, self.moments, eta=0.1)
                if not self.moments or not np.isfinite(self.tc):  # :144-149 (prints, keeps going)
                    if not self.moments:
                        return self
                delta = np.abs(self.tc - last_tc)
                rec["TC"] = float(self.tc)
                self.trace.append(rec)
                self.history["TC"] = self.history.get("TC", []) + [self.moments["TC"]]
                if delta < self.tol:
                    break
        self.moments = self._moments(xt, self.ws, quick=False)
        order = np.argsort(-self.moments["TCs"])
        self.ws = self.ws[order]
        self.moments = self._moments(xt, self.ws, quick=False)
        return self

    def fit_transform(self, x):
        self.fit(x)
        return self.transform(x)

    # -- transform / invert / predict / covariance (:386-395, :431-455) ------------------
    def transform(self, x, details=False):
        xt = self.preprocess(x)
        assert self.nv == xt.shape[1], \
            "Incorrect number of variables in input, %d instead of %d" % (xt.shape[1], self.nv)
        if details:
            return xt.dot(self.ws.T), self._moments(xt, self.ws)
        return xt.dot(self.ws.T)

    def invert(self, x):
        if self.gaussianize == 'standard':
            return self.theta[1] * x + self.theta[0]
        if self.gaussianize == 'outliers':
            return self.theta[1] * unsquash_tails(x) + self.theta[0]
        return x

    def predict(self, y):
        return self.invert(np.dot(self.moments["X_i Z_j"], y.T).T)

    def get_covariance(self):
        m = self.moments
        if self.discourage_overlap:
            z = m['rhoinvrho'] / (1 + m['Si'])
            cov = np.dot(z.T, z)
            cov /= (1. - self.eps ** 2)
        else:
            cov = np.einsum('ij,kj->ik', m["X_i Z_j"], m["X_i Y_j"])
        np.fill_diagonal(cov, 1)
        return self.theta[1][:, np.newaxis] * self.theta[1] * cov


def fit_layers(x, layers, missing_values=None, gaussianize='standard', discourage_overlap=True,
               max_iter=10000, seed=None, work_dtype=np.float32, tol=1e-5):
    """The CLI's hierarchical loop (vis_corex.py:529-545): layer 0 on X (with the missing marker),
    layer k>0 on transform() of layer k-1 (no missing marker).  Appends a final 1-unit layer like
    vis_corex.py:489-491 does.  `seed` is passed to every layer (the CLI itself never seeds)."""
    layers = list(layers)
    if layers[-1] != 1:
        layers.append(1)
    models, x_prev = [], x
    for depth, width in enumerate(layers):
        if depth == 0:
            mdl = OracleCorex(n_hidden=width, gaussianize=gaussianize, missing_values=missing_values,
                              discourage_overlap=discourage_overlap, max_iter=max_iter, seed=seed,
                              work_dtype=work_dtype, tol=tol)
        else:
            x_prev = models[-1].transform(x_prev)
            mdl = OracleCorex(n_hidden=width, gaussianize=gaussianize,
                              discourage_overlap=discourage_overlap, max_iter=max_iter, seed=seed,
                              work_dtype=work_dtype, tol=tol)
        models.append(mdl.fit(x_prev))
    return models


# --------------------------------------------------------------------------------------
# synthetic data: Gaussian latent-factor generator, frozen
# --------------------------------------------------------------------------------------
def latent_factor_data(n_samples, n_vars, n_factors, seed=0, snr=1.0, block=1024, dtype=np.float32,
                       snr_spread=0.0):
    """X[:, i] = (sqrt(snr_g) Z[:, g] + noise) / sqrt(1 + snr_g), g = i mod k, generated in column
    blocks of `block` in index order from RandomState(seed).  `snr_spread` > 0 gives group g the
    signal-to-noise ratio snr * (1 + snr_spread * g) so per-factor TCs are well separated
    (SURVEY.md section 7 hard part 6); 0 reproduces the survey's generator exactly."""
    rng = np.random.RandomState(seed)
    z = rng.randn(n_samples, n_factors)
    x = np.empty((n_samples, n_vars), dtype=dtype)
    for lo in range(0, n_vars, block):
        hi = min(lo + block, n_vars)
        g = np.arange(lo, hi) % n_factors
        s = snr * (1.0 + snr_spread * g)
        noise = rng.randn(n_samples, hi - lo)
        x[:, lo:hi] = ((np.sqrt(s) * z[:, g] + noise) / np.sqrt(1.0 + s)).astype(dtype)
    return x
