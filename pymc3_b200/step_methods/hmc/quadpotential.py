"""Mass-matrix descriptors.  Protocol of pymc3/step_methods/hmc/quadpotential.py:91-132
(`velocity`, `energy`, `velocity_energy`, `random`, `update`, `raise_ok`, `reset`, `.dtype`).

On the sampling path these objects only *describe* the potential (initial mean / variance /
pseudo-sample weight / window); the arithmetic -- v = var (.) p, p ~ N(0,1)/sqrt(var), the
Welford windows of QuadPotentialDiagAdapt (:211-225, :313-353) -- runs per chain on the device
(csrc/b2_core.cuh).  The small NumPy methods below keep the object usable stand-alone, as the
reference's unit tests use it (tests/test_quadpotential.py:48-135); the sampler never calls them.
Dense static potentials (QuadPotentialFull / QuadPotentialFullInv, :400-479) run on the device too:
the engine integrates z = L^-1 q with unit mass, L the Cholesky factor of the covariance (`device_chol`,
b2_set_dense_mass) -- the same flow, U-turn products and energies as the dense metric on q.  The adaptive dense
potential (QuadPotentialFullAdapt, :482-572) and user subclasses are host objects: a step method that holds one is
driven from the host (host_transition.py), logp / dlogp still on the device.
"""
import warnings

import numpy as np
import scipy.linalg

__all__ = ["quad_potential", "QuadPotentialDiag", "QuadPotentialDiagAdapt", "QuadPotentialDiagAdaptGrad",
           "QuadPotentialFull", "QuadPotentialFullInv", "QuadPotentialFullAdapt", "isquadpotential",
           "PositiveDefiniteError"]


class PositiveDefiniteError(ValueError):
    def __init__(self, msg, idx):
        super().__init__(msg)
        self.idx, self.msg = idx, msg

    def __str__(self):
        return "Scaling is not positive definite: %s. Check indexes %s." % (self.msg, self.idx)


def partial_check_positive_definite(C):
    d = C if C.ndim == 1 else np.diag(C)
    bad, = np.nonzero(np.logical_or(np.isnan(d), d <= 0))
    if len(bad):
        raise PositiveDefiniteError("Simple check failed. Diagonal contains negatives", bad)


def quad_potential(C, is_cov):
    """quadpotential.py:30-64: scaling vector -> potential (precision unless is_cov)."""
    C = np.asarray(C, dtype="f8")
    partial_check_positive_definite(C)
    if C.ndim == 1:
        return QuadPotentialDiag(C if is_cov else 1.0 / C)
    return QuadPotentialFull(C) if is_cov else QuadPotentialFullInv(C)


class QuadPotential:
    device_kind = None          # "diag" | "diag_adapt": what the engine can run

    def velocity(self, x, out=None):
        raise NotImplementedError("Abstract method")

    def energy(self, x, velocity=None):
        raise NotImplementedError("Abstract method")

    def random(self):
        raise NotImplementedError("Abstract method")

    def velocity_energy(self, x, v_out):
        raise NotImplementedError("Abstract method")

    def update(self, sample, grad, tune):
        pass

    def raise_ok(self, vmap=None):
        return None

    def reset(self):
        pass


def isquadpotential(value):
    return isinstance(value, QuadPotential)


class _DiagMath(QuadPotential):
    def velocity(self, x, out=None):
        return np.multiply(self._var, x, out=out)

    def energy(self, x, velocity=None):
        if velocity is None:
            velocity = self._var * x
        return 0.5 * np.dot(x, velocity)

    def velocity_energy(self, x, v_out):
        np.multiply(self._var, x, out=v_out)
        return 0.5 * np.dot(x, v_out)

    def random(self):
        return (np.random.normal(size=self._n) / np.sqrt(self._var)).astype(self.dtype)


class QuadPotentialDiag(_DiagMath):
    """Static diagonal potential; `v` is the covariance diagonal (quadpotential.py:356-397)."""

    device_kind = "diag"

    def __init__(self, v, dtype=None):
        self.dtype = np.dtype(dtype or "float64")
        self._var = np.asarray(v, dtype=self.dtype).copy()
        self._n = len(self._var)
        self.v = self._var
        self.s = self._var ** 0.5
        self.inv_s = 1.0 / self.s

    def device_init(self):
        return dict(mean=np.zeros(self._n), var=self._var.astype("f8"), weight=0.0, window=101, adapt=0)


class QuadPotentialDiagAdapt(_DiagMath):
    """Adaptive diagonal potential (quadpotential.py:140-269)."""

    device_kind = "diag_adapt"

    def __init__(self, n, initial_mean, initial_diag=None, initial_weight=0, adaptation_window=101,
                 adaptation_window_multiplier=1, dtype=None):
        initial_mean = np.asarray(initial_mean)
        if initial_diag is not None and np.ndim(initial_diag) != 1:
            raise ValueError("Initial diagonal must be one-dimensional.")
        if initial_mean.ndim != 1:
            raise ValueError("Initial mean must be one-dimensional.")
        if initial_diag is not None and len(initial_diag) != n:
            raise ValueError("Wrong shape for initial_diag: expected %s got %s" % (n, len(initial_diag)))
        if len(initial_mean) != n:
            raise ValueError("Wrong shape for initial_mean: expected %s got %s" % (n, len(initial_mean)))
        if adaptation_window_multiplier != 1:
            raise NotImplementedError("adaptation_window_multiplier != 1 is not supported on the device")
        self.dtype = np.dtype(dtype or "float64")
        if initial_diag is None:
            initial_diag = np.ones(n, dtype=self.dtype)
            initial_weight = 1
        self._n = n
        self._initial_mean = np.array(initial_mean, dtype="f8")
        self._initial_diag = np.array(initial_diag, dtype="f8")
        self._initial_weight = float(initial_weight)
        self.adaptation_window = int(adaptation_window)
        self.reset()

    def reset(self):
        self._var = self._initial_diag.astype(self.dtype).copy()
        self._stds = np.sqrt(self._var)
        # host-side Welford windows: used only when a user SUBCLASS of this potential drives host transitions
        # (quadpotential.py:211-225, 313-353); the batched path keeps these windows on the device
        self._fg = [self._initial_weight, self._initial_mean.copy(), self._initial_diag * self._initial_weight]
        self._bg = [0.0, np.zeros(self._n), np.zeros(self._n)]
        self._n_samples = 0

    @staticmethod
    def _welford_add(win, x):
        win[0] += 1.0
        old = x - win[1]
        win[1] = win[1] + old / win[0]
        win[2] = win[2] + old * (x - win[1])

    def update(self, sample, grad, tune):
        if not tune:
            return
        x = np.asarray(sample, dtype="f8")
        window = self.adaptation_window
        self._welford_add(self._fg, x)
        self._welford_add(self._bg, x)
        self._var = (self._fg[2] / self._fg[0]).astype(self.dtype)
        self._stds = np.sqrt(self._var)
        if self._n_samples > 0 and self._n_samples % window == 0:
            self._fg = self._bg
            self._bg = [0.0, np.zeros(self._n), np.zeros(self._n)]
        self._n_samples += 1

    def device_init(self):
        return dict(mean=self._initial_mean, var=self._initial_diag, weight=self._initial_weight,
                    window=self.adaptation_window, adapt=1)

    def sync(self, var):
        """Adopt one chain's adapted variances from the device (step.potential._var inspection)."""
        self._var = np.asarray(var, dtype=self.dtype).copy()
        self._stds = np.sqrt(self._var)

    def raise_ok(self, vmap):
        """quadpotential.py:227-269: name the RV whose mass-matrix entry is zero / non-finite."""
        for bad, what, tail in ((self._stds == 0, "zeros", "zero"),
                                (~np.isfinite(self._stds), "non-finite values", "non-finite")):
            if not np.any(bad):
                continue
            names = []
            for vm in vmap:
                names.extend((vm.var, i) for i in range(vm.slc.stop - vm.slc.start))
            lines = ["Mass matrix contains %s on the diagonal. " % what]
            for ii in np.where(bad)[0]:
                lines.append("The derivative of RV `{}`.ravel()[{}] is {}.".format(names[ii][0], names[ii][1], tail))
            raise ValueError("\n".join(lines))


class _DenseMath(QuadPotential):
    """energy and velocity_energy in terms of velocity (quadpotential.py:425-436, 463-472)"""

    def energy(self, x, velocity=None):
        if velocity is None:
            velocity = self.velocity(x)
        return 0.5 * np.dot(x, velocity)

    def velocity_energy(self, x, v_out):
        self.velocity(x, out=v_out)
        return 0.5 * np.dot(x, v_out)


class _DenseDevice:
    device_kind = "dense"

    def device_init(self):
        n = self._n
        return dict(mean=np.zeros(n), var=np.ones(n), weight=0.0, window=101, adapt=0)


class QuadPotentialFull(_DenseDevice, _DenseMath):
    """Dense potential given the covariance (quadpotential.py:438-472): v = cov p, p ~ chol^-T z."""

    def __init__(self, cov, dtype=None):
        self.dtype = np.dtype(dtype or "float64")
        self._cov = np.array(cov, dtype=self.dtype, copy=True)
        self._chol = scipy.linalg.cholesky(self._cov, lower=True)
        self._n = len(self._cov)

    def velocity(self, x, out=None):
        return np.dot(self._cov, x, out=out)

    def random(self):
        z = np.random.normal(size=self._n).astype(self.dtype)
        return scipy.linalg.solve_triangular(self._chol.T, z, overwrite_b=True)

    def device_chol(self):
        return np.asarray(self._chol, dtype="f8")


class QuadPotentialFullInv(_DenseDevice, _DenseMath):
    """Dense potential given the precision A (quadpotential.py:400-436): v = A^-1 p via Cholesky, p = L z."""

    def __init__(self, A, dtype=None):
        self.dtype = np.dtype(dtype or "float64")
        self.L = scipy.linalg.cholesky(np.asarray(A, dtype=self.dtype), lower=True)
        self._n = self.L.shape[0]

    def device_chol(self):
        """lower factor of the covariance A^-1 = (L L^T)^-1"""
        eye = np.eye(self._n)
        cov = scipy.linalg.cho_solve((np.asarray(self.L, dtype="f8"), True), eye)
        return scipy.linalg.cholesky(0.5 * (cov + cov.T), lower=True)

    def velocity(self, x, out=None):
        vel = scipy.linalg.cho_solve((self.L, True), x)
        if out is None:
            return vel
        out[:] = vel
        return out

    def random(self):
        return np.dot(self.L, np.random.normal(size=self.L.shape[0]).astype(self.dtype))


class QuadPotentialDiagAdaptGrad(QuadPotentialDiagAdapt):
    """Diagonal potential adapted from the mean absolute GRADIENT after the first 150 tuning draws
    (quadpotential.py:272-310, experimental there too).  A host object: the step method that holds it is driven
    from the host (host_transition.py), so `update` sees every (sample, grad)."""

    device_kind = None

    def reset(self):
        super().reset()
        self._abs_grad = [np.zeros(self._n), np.zeros(self._n)]      # running sums of |grad|: in use / next
        self._n_grad = [0, 0]

    def update(self, sample, grad, tune):
        if not tune:
            return
        g = np.abs(np.asarray(grad, dtype="f8"))
        for w in (0, 1):
            self._abs_grad[w] += g
            self._n_grad[w] += 1
        if self._n_samples <= 150:
            super().update(sample, grad, tune)                       # sample variances first (counts the draw itself)
        else:
            self._var = ((self._n_grad[0] / self._abs_grad[0]) ** 2).astype(self.dtype)
            self._stds = np.sqrt(self._var)
            self._n_samples += 1
        if self._n_samples > 100 and self._n_samples % 100 == 50:    # the younger window takes over, a fresh one starts
            self._abs_grad = [self._abs_grad[1], np.ones(self._n)]
            self._n_grad = [self._n_grad[1], 1]


class _RunningCovariance:
    """Welford mean / co-moment accumulator seeded with `weight` pseudo-observations (quadpotential.py:575-628)."""

    def __init__(self, n, mean=None, cov=None, weight=0.0):
        self.count = float(weight)
        self.mean = np.zeros(n) if mean is None else np.array(mean, dtype="f8")
        self.comoment = (np.eye(n) if cov is None else np.array(cov, dtype="f8")) * self.count
        if self.comoment.shape != (n, n):
            raise ValueError("Invalid shape for initial covariance.")
        if self.mean.shape != (n,):
            raise ValueError("Invalid shape for initial mean.")

    def add_sample(self, x, weight=1.0):
        x = np.asarray(x, dtype="f8")
        self.count += 1.0
        before = x - self.mean
        self.mean = self.mean + before / self.count
        self.comoment = self.comoment + weight * np.outer(x - self.mean, before)

    def current_covariance(self):
        if self.count == 0:
            raise ValueError("Can not compute covariance without samples.")
        return self.comoment / (self.count - 1.0)

    def current_mean(self):
        return self.mean.copy()


_WeightedCovariance = _RunningCovariance      # the reference's name for it


class QuadPotentialFullAdapt(QuadPotentialFull):
    """Dense potential adapted to the sample covariance in growing windows (quadpotential.py:482-572; experimental
    there too).  A host object like every adaptive / user potential beyond the diagonal ones."""

    device_kind = None

    def __init__(self, n, initial_mean, initial_cov=None, initial_weight=0, adaptation_window=101,
                 adaptation_window_multiplier=2, update_window=1, dtype=None):
        warnings.warn("QuadPotentialFullAdapt is an experimental feature")
        initial_mean = np.asarray(initial_mean)
        if initial_cov is not None and np.ndim(initial_cov) != 2:
            raise ValueError("Initial covariance must be two-dimensional.")
        if initial_mean.ndim != 1:
            raise ValueError("Initial mean must be one-dimensional.")
        if initial_cov is not None and np.shape(initial_cov) != (n, n):
            raise ValueError("Wrong shape for initial_cov: expected %s got %s" % (n, np.shape(initial_cov)))
        if len(initial_mean) != n:
            raise ValueError("Wrong shape for initial_mean: expected %s got %s" % (n, len(initial_mean)))
        if initial_cov is None:
            initial_cov, initial_weight = np.eye(n), 1
        self._args = (n, np.array(initial_mean, dtype="f8"), np.array(initial_cov, dtype="f8"), float(initial_weight),
                      int(adaptation_window))
        self.adaptation_window_multiplier = float(adaptation_window_multiplier)
        self._update_window = int(update_window)
        super().__init__(initial_cov, dtype=dtype)
        self.reset()

    def reset(self):
        n, mean, cov, weight, window = self._args
        self._cov = cov.astype(self.dtype)
        self._chol = scipy.linalg.cholesky(self._cov, lower=True)
        self._chol_error = None
        self._fg = _RunningCovariance(n, mean, cov, weight)
        self._bg = _RunningCovariance(n)
        self._n_samples = 0
        self._previous_update = 0
        self.adaptation_window = window

    def update(self, sample, grad, tune):
        if not tune:
            return
        since = self._n_samples - self._previous_update
        self._fg.add_sample(sample)
        self._bg.add_sample(sample)
        if (since + 1) % self._update_window == 0:
            self._cov = self._fg.current_covariance().astype(self.dtype)
            try:
                self._chol = scipy.linalg.cholesky(self._cov, lower=True)
            except (scipy.linalg.LinAlgError, ValueError) as error:
                self._chol_error = error
        if since >= self.adaptation_window:                          # the background window becomes the estimate
            self._fg, self._bg = self._bg, _RunningCovariance(self._n)
            self._previous_update = self._n_samples
            self.adaptation_window = int(self.adaptation_window * self.adaptation_window_multiplier)
        self._n_samples += 1

    def raise_ok(self, vmap=None):
        if self._chol_error is not None:
            raise ValueError(str(self._chol_error))
