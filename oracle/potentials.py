"""Mass-matrix (kinetic-energy) objects of the oracle.  TEST INFRASTRUCTURE.

Restates ``pymc3/step_methods/hmc/quadpotential.py`` (diagonal family only):
``_WeightedVariance`` :313-353, ``QuadPotentialDiagAdapt`` :140-269,
``QuadPotentialDiag`` :356-397, ``quad_potential`` :30-64 and
``tuning/scaling.py:80-102`` (``guess_scaling``'s clipping).
All arithmetic is float64 (the reference's default floatX).
"""
import numpy as np


class RunningVariance:
    """Welford mean/variance with pseudo-sample seeding (quadpotential.py:313-353)."""

    def __init__(self, n, mean0=None, var0=None, weight0=0):
        self.count = float(weight0)                              # :318
        self.mean = np.zeros(n) if mean0 is None else np.array(mean0, dtype="d")
        self.m2 = np.zeros(n) if var0 is None else np.array(var0, dtype="d")
        self.m2 *= self.count                                    # :329

    def push(self, x):
        # :336-342 (weight is always 1 on the sampler path)
        x = np.asarray(x, dtype="d")
        self.count += 1
        before = x - self.mean
        self.mean += before / self.count
        after = x - self.mean
        self.m2 += before * after

    def variance(self):
        if self.count == 0:
            raise ValueError("Can not compute variance without samples.")   # :345-346
        return self.m2 / self.count                              # population variance :347-350


class DiagAdaptPotential:
    """quadpotential.py:140-269  (QuadPotentialDiagAdapt)."""

    adaptive = True

    def __init__(self, n, initial_mean, initial_diag=None, initial_weight=0,
                 adaptation_window=101):
        initial_mean = np.asarray(initial_mean, dtype="d")
        if initial_diag is None:                                  # :166-168
            initial_diag = np.ones(n)
            initial_weight = 1
        initial_diag = np.asarray(initial_diag, dtype="d")
        if initial_diag.ndim != 1 or initial_mean.ndim != 1:
            raise ValueError("Initial mean/diagonal must be one-dimensional.")
        if len(initial_diag) != n or len(initial_mean) != n:
            raise ValueError("Wrong shape for initial mean/diag")
        self.n = n
        self.var = initial_diag.copy()
        self.stds = np.sqrt(self.var)
        self.inv_stds = 1.0 / self.stds
        self.fg = RunningVariance(n, initial_mean, initial_diag, initial_weight)   # :177-178
        self.bg = RunningVariance(n)                                               # :179
        self.n_seen = 0
        self.window = adaptation_window

    def velocity(self, p):
        return self.var * p                                       # :185-187

    def kinetic(self, p, v):
        return 0.5 * np.dot(p, v)                                 # :189-198

    def draw(self, normals):
        return self.inv_stds * normals                            # :200-203

    def update(self, q, tune):
        # :211-225 -- refresh var from the foreground window after EVERY tuning draw;
        # swap windows when n_seen is a positive multiple of the window; the counter is
        # bumped after that test, so the first swap happens on the 102nd tuning draw.
        if not tune:
            return
        self.fg.push(q)
        self.bg.push(q)
        self.var = self.fg.variance()
        self.stds = np.sqrt(self.var)
        self.inv_stds = 1.0 / self.stds
        if self.n_seen > 0 and self.n_seen % self.window == 0:
            self.fg = self.bg
            self.bg = RunningVariance(self.n)
        self.n_seen += 1

    def check(self):
        # :243-269 (message shape only; RV names are added by the host layer)
        if np.any(self.stds == 0):
            raise ValueError("Mass matrix contains zeros on the diagonal. ")
        if np.any(~np.isfinite(self.stds)):
            raise ValueError("Mass matrix contains non-finite values on the diagonal. ")


class DiagPotential:
    """quadpotential.py:356-397  (QuadPotentialDiag): v is the covariance diagonal."""

    adaptive = False

    def __init__(self, v):
        self.var = np.asarray(v, dtype="d").copy()
        self.stds = self.var ** 0.5
        self.inv_stds = 1.0 / self.stds
        self.n = len(self.var)

    def velocity(self, p):
        return self.var * p

    def kinetic(self, p, v):
        return 0.5 * np.dot(p, v)

    def draw(self, normals):
        return normals * self.inv_stds                            # :378-380

    def update(self, q, tune):
        pass

    def check(self):
        pass


def potential_from_scaling(c, is_cov):
    """quadpotential.py:30-64 restricted to 1-d scalings (precision unless is_cov)."""
    c = np.asarray(c, dtype="d")
    if c.ndim != 1:
        raise NotImplementedError("dense mass matrices are a 'next' row (SURVEY 8f N2)")
    if np.any(np.isnan(c) | (c <= 0)):
        raise ValueError("Scaling is not positive definite: Simple check failed. "
                         "Diagonal contains negatives")
    return DiagPotential(c if is_cov else 1.0 / c)


def clip_precision(tau, bound=1e-8):
    """tuning/scaling.py:98-102 (adjust_precision)."""
    mag = np.sqrt(np.abs(tau))
    lg = np.clip(np.log(mag), np.log(bound), np.log(1.0 / bound))
    return np.exp(lg) ** 2
