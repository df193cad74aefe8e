"""Model objects and the logp/dlogp handle of the B200 engine.

The reference turns a Theano graph into one compiled function per model
(`Model.logp_dlogp_function` pymc3/model.py:885 -> `ValueGradFunction` :541-713).  Here a model
is one of the fused CUDA families; the objects below keep the reference's *interface*:

* `Model` is a context manager (`with model:` like pymc3/model.py:170-230 `Context`), exposes
  `test_point`, `ndim`, `free_RVs`/`unobserved_RVs` names, `dict_to_array`, `logp_dlogp_function`.
* `ValueGradFunction` mirrors model.py:633-700: `__call__(array, grad_out=None)`, `size`, `dtype`,
  `_ordering.vmap`, `set_extra_values`, `dict_to_array`, `array_to_dict`, `array_to_full_dict`.
* `ArrayOrdering` / `VarMap` mirror pymc3/blocking.py:26-59 (flat slices in creation order).
* transformed names follow pymc3/util.py:50-66 (`<name>_log__`).
"""
import collections
import json
import os
import threading

import numpy as np

from . import _capi

VarMap = collections.namedtuple("VarMap", "var, slc, shp, dtyp")
_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


class ArrayOrdering:
    """pymc3/blocking.py:33-59: an ordering for an array space."""

    def __init__(self, names_shapes, dtype):
        self.vmap, self.by_name, off = [], {}, 0
        for name, shape in names_shapes:
            n = int(np.prod(shape)) if shape else 1
            vm = VarMap(name, slice(off, off + n), tuple(shape), np.dtype(dtype))
            self.vmap.append(vm)
            self.by_name[name] = vm
            off += n
        self.size = off

    def __getitem__(self, key):
        return self.by_name[key]


class _Context:
    _stack = threading.local()

    @classmethod
    def stack(cls):
        if not hasattr(cls._stack, "items"):
            cls._stack.items = []
        return cls._stack.items


def modelcontext(model=None):
    """pymc3/model.py:232-245."""
    if model is not None:
        return model
    st = _Context.stack()
    if not st:
        raise TypeError("No model on context stack.")
    return st[-1]


class Model:
    """Base class of the fused model families (the engine's stand-in for pm.Model)."""

    family = None
    free = ()                  # ((name, shape), ...) creation order, transformed names
    log_transformed = {}       # deterministic name -> free (log) name

    def __enter__(self):
        _Context.stack().append(self)
        return self

    def __exit__(self, *exc):
        _Context.stack().pop()

    # -- naming / layout
    @property
    def ndim(self):
        return self.ordering("float64").size

    def ordering(self, dtype="float64"):
        return ArrayOrdering(self.free, dtype)

    @property
    def free_RVs(self):
        return [n for n, _ in self.free]

    @property
    def cont_vars(self):
        return self.free_RVs

    @property
    def vars(self):
        return self.free_RVs

    @property
    def deterministics(self):
        return list(self.log_transformed.keys())

    @property
    def unobserved_RVs(self):          # model.py:954-957: free (transformed) + deterministics
        return self.free_RVs + self.deterministics

    def _test_values(self):
        """free-variable name -> default test value where it is not 0 (the reference takes each distribution's
        first available default of ('median', 'mean', 'mode'), distribution.py:208, through its transform:
        HalfCauchy -> log(beta) continuous.py:2401-2402, Exponential -> log(log 2 / lam) :1519-1521)."""
        return {}

    @property
    def test_point(self):
        tv = self._test_values()
        return {n: np.full(s, float(tv.get(n, 0.0))) for n, s in self.free}

    def dict_to_array(self, point):
        od = self.ordering()
        out = np.empty(od.size)
        for vm in od.vmap:
            out[vm.slc] = np.ravel(point[vm.var])
        return out

    def array_to_dict(self, array):
        od = self.ordering()
        array = np.asarray(array)
        return {vm.var: array[vm.slc].reshape(vm.shp) for vm in od.vmap}

    def expand(self, qs):
        """[..., D] flat positions -> {name: [..., *shape]} for all unobserved RVs (vectorised
        replacement of the per-draw `fastfn` call in backends/ndarray.py:266)."""
        qs = np.asarray(qs)
        out = {}
        for vm in self.ordering().vmap:
            out[vm.var] = qs[..., vm.slc].reshape(qs.shape[:-1] + vm.shp)
        for det, src in self.log_transformed.items():
            out[det] = np.exp(out[src])
        return out

    def logp_dlogp_function(self, grad_vars=None, dtype="float64", device=0, **kwargs):
        return ValueGradFunction(self, dtype=dtype, device=device)

    def engine(self, n_chains, dtype="float32", device=0):
        from .engine import Engine
        return Engine(self._describe, n_chains, dtype=dtype, device=device)

    def _describe(self, upload):
        raise NotImplementedError


class ValueGradFunction:
    """Device-backed mirror of pymc3/model.py:541-713 for one point at a time."""

    def __init__(self, model, dtype="float64", device=0, batch=1):
        self._model = model
        self.dtype = np.dtype(dtype)
        if self.dtype not in (np.dtype("float32"), np.dtype("float64")):
            raise TypeError("Invalid dtype. Must be a floating point type.")      # model.py:596-598
        self._ordering = model.ordering(self.dtype)
        self.size = self._ordering.size
        self._engine = model.engine(batch, dtype=self.dtype.name, device=device)
        self._extra_are_set = True

    def set_extra_values(self, extra_vars):
        self._extra_are_set = True           # the fused families have no extra (non-grad) variables

    def get_extra_values(self):
        return {}

    def __call__(self, array, grad_out=None, extra_vars=None):
        array = np.asarray(array)
        if array.shape != (self.size,):
            raise ValueError("Invalid shape for array. Must be %s but is %s." % ((self.size,), array.shape))
        logp, grad = self._engine.logp_dlogp(array.reshape(1, -1))
        logp = float(logp.cpu().numpy()[0])
        grad = grad.cpu().numpy()[0]
        if grad_out is None:
            return np.array(logp), grad.astype(self.dtype)
        np.copyto(grad_out, grad)
        return np.array(logp)

    def batch(self, arrays):
        """[n, D] -> (logp [n], grad [n, D]) evaluated in one launch."""
        arrays = np.asarray(arrays)
        if self._engine.n_chains < len(arrays):
            self._engine = self._model.engine(len(arrays), dtype=self.dtype.name, device=self._engine.device)
        logp, grad = self._engine.logp_dlogp(arrays)
        return logp.cpu().numpy(), grad.cpu().numpy()

    def dict_to_array(self, point):
        return self._model.dict_to_array(point).astype(self.dtype)

    def array_to_dict(self, array):
        if array.shape != (self.size,):
            raise ValueError("Array should have shape (%s,) but has %s" % (self.size, array.shape))
        if array.dtype != self.dtype:
            raise ValueError("Array has invalid dtype. Should be %s but is %s" % (self.dtype, array.dtype))
        return self._model.array_to_dict(array)

    def array_to_full_dict(self, array):
        return self.array_to_dict(array)

    def profile(self, array, *args, **kwargs):
        raise NotImplementedError("use ncu; see profiles/ (replaces theano profiling, model.py:668-671)")


# --------------------------------------------------------------------------- model families
class StdNormal(Model):
    """x[n] ~ Normal(mu, sigma): the benchmarks.py:75-91 overhead model / posterior fixtures."""

    family = _capi.B2_STD_NORMAL

    def __init__(self, n=1, mu=0.0, sigma=1.0, name="x"):
        self.n = int(n)
        self.mu = np.broadcast_to(np.asarray(mu, dtype="f8"), (self.n,)).copy()
        self.sigma = np.broadcast_to(np.asarray(sigma, dtype="f8"), (self.n,)).copy()
        self.free = ((name, (self.n,)),)

    def _describe(self, upload):
        d = _capi.ModelDesc(family=self.family, D=self.n, N=self.n, G=0)
        d.d_aux0 = upload(self.mu, "f8")
        d.d_aux1 = upload(self.sigma, "f8")
        return d


class EightSchoolsNCP(Model):
    """pymc3/examples/gelman_schools.py:26-40 (config C1)."""

    family = _capi.B2_EIGHT_SCHOOLS_NCP
    log_transformed = {"tau": "tau_log__"}

    def __init__(self, y=None, sigma=None, mu_sd=1e6, tau_beta=25.0):
        self.y = np.array([28, 8, -3, 7, -1, 1, 18, 12], dtype="f8") if y is None else np.asarray(y, "f8")
        self.sigma = (np.array([15, 10, 16, 11, 9, 11, 10, 18], dtype="f8") if sigma is None
                      else np.asarray(sigma, "f8"))
        self.J = len(self.y)
        self.mu_sd, self.tau_beta = float(mu_sd), float(tau_beta)
        self.free = (("eta", (self.J,)), ("mu", ()), ("tau_log__", ()))

    def _test_values(self):
        return {"tau_log__": np.log(self.tau_beta)}

    def _describe(self, upload):
        d = _capi.ModelDesc(family=self.family, D=self.J + 2, N=self.J, G=0)
        d.d_aux0 = upload(self.y, "f8")
        d.d_aux1 = upload(self.sigma, "f8")
        d.hp[0], d.hp[1] = self.mu_sd, self.tau_beta
        return d


class LogisticGLM(Model):
    """pm.glm.GLM(..., family=Binomial()) of glm/linear.py:49-101 + glm/families.py:115-119 (C2, C5).

    One scalar RV per column as the reference creates them: Intercept ~ Flat,
    x_k ~ Normal(0, tau=1e-6).
    """

    family = _capi.B2_GLM_LOGIT

    def __init__(self, X, y, labels=None, prior_tau=1e-6):
        if type(X).__module__.startswith("torch"):     # device-resident data (large, generated on the GPU)
            self.X, self.y = X, y
        else:
            self.X = np.ascontiguousarray(X, dtype="f4")
            self.y = np.ascontiguousarray(y, dtype="f4")
        if self.y.ndim > 1:
            raise TypeError("Only one-dimensional observed variable objects (i.e. of shape `(n, )`) "
                            "are supported")                       # glm/linear.py:55-58
        self.N, self.K = self.X.shape
        if labels is None:
            labels = ["x%d" % i for i in range(self.K)]
        self.prior_tau = float(prior_tau)
        self.free = tuple((n, ()) for n in ["Intercept"] + list(labels))

    def _describe(self, upload):
        d = _capi.ModelDesc(family=self.family, D=self.K + 1, N=self.N, G=self.K)
        d.d_X = upload(self.X, "f4")
        d.d_y = upload(self.y, "f4")
        d.hp[0] = self.prior_tau
        return d


class HierLinearNCP(Model):
    """benchmarks/benchmarks/benchmarks.py:25-45 radon model, non-centred (C3)."""

    family = _capi.B2_HIER_LINEAR_NCP
    log_transformed = {"sigma_a": "sigma_a_log__", "sigma_b": "sigma_b_log__", "eps": "eps_log__"}

    def __init__(self, group_idx, floor, y, n_groups=None, mu_sd=100.0 ** 2, hc_beta=5.0):
        idx = np.asarray(group_idx, dtype=np.int64)
        self.G = int(n_groups if n_groups is not None else idx.max() + 1)
        order = np.argsort(idx, kind="stable")          # group-sorted layout (csrc/b2_hier.cu)
        self.idx_sorted = idx[order]
        self.floor = np.asarray(floor)[order].astype(np.uint8)
        self.y = np.asarray(y, dtype="f4")[order]
        self.N = len(self.y)
        self.grp_off = np.concatenate([[0], np.cumsum(np.bincount(idx, minlength=self.G))]).astype(np.int32)
        self.mu_sd, self.hc_beta = float(mu_sd), float(hc_beta)
        G = self.G
        self.free = (("mu_a", ()), ("sigma_a_log__", ()), ("mu_b", ()), ("sigma_b_log__", ()),
                     ("a", (G,)), ("b", (G,)), ("eps_log__", ()))

    def _test_values(self):
        v = np.log(self.hc_beta)
        return {"sigma_a_log__": v, "sigma_b_log__": v, "eps_log__": v}

    def _describe(self, upload):
        d = _capi.ModelDesc(family=self.family, D=2 * self.G + 5, N=self.N, G=self.G)
        d.d_y = upload(self.y, "f4")
        d.d_floor = upload(self.floor, np.uint8)
        d.d_grp_off = upload(self.grp_off, np.int32)
        d.hp[0], d.hp[1] = self.mu_sd, self.hc_beta
        return d


class StochVol(Model):
    """docs/source/notebooks/stochastic_volatility.ipynb cell 10 (C4)."""

    family = _capi.B2_STOCH_VOL
    log_transformed = {"step_size": "step_size_log__", "nu": "nu_log__"}

    def __init__(self, returns=None, step_lam=10.0, nu_lam=0.1):
        if returns is None:
            returns = sp500_log_returns()
        self.returns = np.asarray(returns, dtype="f8")
        self.T = len(self.returns)
        self.step_lam, self.nu_lam = float(step_lam), float(nu_lam)
        self.free = (("step_size_log__", ()), ("volatility", (self.T,)), ("nu_log__", ()))

    def _test_values(self):
        return {"step_size_log__": np.log(np.log(2.0) / self.step_lam), "nu_log__": np.log(np.log(2.0) / self.nu_lam)}

    def _describe(self, upload):
        d = _capi.ModelDesc(family=self.family, D=self.T + 2, N=self.T, G=0)
        d.d_aux0 = upload(self.returns, "f8")
        d.hp[0], d.hp[1] = self.step_lam, self.nu_lam
        return d


# ---------------------------------------------------------------------------------- data
def sp500_log_returns():
    """Log-returns of the reference's examples/data/SP500.csv (fixture, tests/golden/make_golden.py)."""
    return np.load(os.path.join(_DATA, "sp500_log_returns.npy"))


def radon_county_counts():
    with open(os.path.join(_DATA, "radon_county_counts.json")) as f:
        return json.load(f)
