"""Mass-matrix descriptors.  Protocol of pymc3/step_methods/hmc/quadpotential.py:91-132
(`velocity`, `energy`, `velocity_energy`, `random`, `update`, `raise_ok`, `reset`, `.dtype`).

On the sampling path these objects only *describe* the potential (initial mean / variance /
pseudo-sample weight / window); the arithmetic -- v = var (.) p, p ~ N(0,1)/sqrt(var), the
Welford windows of QuadPotentialDiagAdapt (:211-225, :313-353) -- runs per chain on the device
(csrc/b2_core.cuh).  The small NumPy methods below keep the object usable stand-alone, as the
reference's unit tests use it (tests/test_quadpotential.py:48-135); the sampler never calls them.
Dense potentials (QuadPotentialFull / QuadPotentialFullInv, :400-470) are host objects: a step
method that holds one is driven from the host (host_transition.py) -- a [D, D] matrix-vector product per leapfrog on
the host, logp / dlogp on the device.  They are not part of the chain-batched device path.
"""
import numpy as np
import scipy.linalg

__all__ = ["quad_potential", "QuadPotentialDiag", "QuadPotentialDiagAdapt", "QuadPotentialFull",
           "QuadPotentialFullInv", "isquadpotential", "PositiveDefiniteError"]


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


class QuadPotentialFull(_DenseMath):
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


class QuadPotentialFullInv(_DenseMath):
    """Dense potential given the precision A (quadpotential.py:400-436): v = A^-1 p via Cholesky, p = L z."""

    def __init__(self, A, dtype=None):
        self.dtype = np.dtype(dtype or "float64")
        self.L = scipy.linalg.cholesky(np.asarray(A, dtype=self.dtype), lower=True)

    def velocity(self, x, out=None):
        vel = scipy.linalg.cho_solve((self.L, True), x)
        if out is None:
            return vel
        out[:] = vel
        return out

    def random(self):
        return np.dot(self.L, np.random.normal(size=self.L.shape[0]).astype(self.dtype))
