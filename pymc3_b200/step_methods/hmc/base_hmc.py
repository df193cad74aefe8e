"""BaseHMC step-method interface over the device engine.

Constructor arguments, attributes and the `step(point) -> (point, [stats])` contract follow
pymc3/step_methods/hmc/base_hmc.py:36-235 and arraystep.py:236-269 (GradientSharedStep).
What differs is where the work happens: the transition (`astep`, base_hmc.py:133-199) runs in
the CUDA state machine for *all chains at once* when driven by `pymc3_b200.sample`
(`_batched = True` advertises that); `step(point)` drives the same device code
for one chain and one draw, for `iter_sample`, user loops and the API tests.
"""
import logging

import numpy as np

from ... import _capi
from ...backends.report import SamplerWarning, WarningType
from ...exceptions import SamplingError
from ...model import modelcontext
from .. import step_sizes
from .quadpotential import (QuadPotentialDiag, QuadPotentialDiagAdapt, QuadPotentialFull, QuadPotentialFullInv,
                            quad_potential)

logger = logging.getLogger("pymc3")


def guess_scaling(point, model, logp_dlogp, scaling_bound=1e-8):
    """pymc3/tuning/scaling.py:80-102 with the diagonal Hessian taken from central differences
    of the device gradient (one batched launch) instead of theanof.hessian_diag."""
    q = model.dict_to_array(point)
    D = len(q)
    h = 1e-4 * np.maximum(1.0, np.abs(q))
    pts = np.concatenate([q + np.diag(h), q - np.diag(h)])
    _, g = logp_dlogp.batch(pts)
    diag = -(g[np.arange(D), np.arange(D)] - g[D + np.arange(D), np.arange(D)]) / (2 * h)
    mag = np.sqrt(np.abs(diag))
    bounded = np.clip(np.log(mag), np.log(scaling_bound), np.log(1.0 / scaling_bound))
    return np.exp(bounded) ** 2


class BaseHMC:
    """Superclass of the Hamiltonian samplers."""

    default_blocked = True
    generates_stats = True
    _kind = None

    def __init__(self, vars=None, scaling=None, step_scale=0.25, is_cov=False, model=None, blocked=True,
                 potential=None, dtype=None, Emax=1000, target_accept=0.8, gamma=0.05, k=0.75, t0=10,
                 adapt_step_size=True, step_rand=None, device=0, exec_mode="auto", glm_path="auto", **kwargs):
        if kwargs:
            raise ValueError("Unknown arguments: %s" % sorted(kwargs))
        self._model = model = modelcontext(model)
        if vars is not None and list(map(str, vars)) != list(model.free_RVs):
            raise NotImplementedError("the device engine samples all continuous variables of the model "
                                      "jointly (blocked); got vars=%s" % (vars,))
        self.vars = list(model.free_RVs)
        self.blocked = blocked
        self.dtype = np.dtype(dtype or "float32")
        if self.dtype not in (np.dtype("float32"), np.dtype("float64")):
            raise TypeError("Invalid dtype %s" % dtype)
        self.device = device
        self.adapt_step_size = adapt_step_size
        self.Emax = Emax
        self.iter_count = 0
        self._ordering = model.ordering(self.dtype)
        size = self._ordering.size
        self.step_size = step_scale / (size ** 0.25)                 # base_hmc.py:93
        self._initial_step = self.step_size
        self.target_accept = target_accept
        self._da = (gamma, k, t0)
        self.step_adapt = step_sizes.DualAverageAdaptation(self.step_size, target_accept, gamma, k, t0)
        self.tune = True
        self.__logp_dlogp_func = None

        if scaling is None and potential is None:                    # base_hmc.py:100-103
            potential = QuadPotentialDiagAdapt(size, np.zeros(size), np.ones(size), 10)
        if isinstance(scaling, dict):                                # base_hmc.py:105-107
            scaling = guess_scaling(scaling, model, self._logp_dlogp_func)
        if scaling is not None and potential is not None:
            raise ValueError("Can not specify both potential and scaling.")
        self.potential = potential if potential is not None else quad_potential(scaling, is_cov)
        # The built-in diagonal potentials are run by the device state machine.  Anything else that speaks the
        # QuadPotential protocol (a user subclass overriding velocity / energy / random / update, test_quadpotential.py:
        # 138-155; a custom step_rand callable) is honoured by driving the tree from the host and calling the user's
        # object per leapfrog, with logp / dlogp still evaluated on the device (host_transition.py).
        # Exactly QuadPotentialFull / FullInv (static dense) also run on the device, in z = L^-1 q (engine.set_dense_mass).
        self._dense = type(self.potential) in (QuadPotentialFull, QuadPotentialFullInv) and size <= 1024
        self._host_driven = type(self.potential) not in (QuadPotentialDiag, QuadPotentialDiagAdapt) and not self._dense
        if self._host_driven and not all(hasattr(self.potential, m) for m in ("velocity", "energy", "random")):
            raise TypeError("potential must implement the QuadPotential protocol (velocity, energy, random); "
                            "got %r" % type(self.potential).__name__)
        if step_rand is not None and not getattr(step_rand, "_b2_unif", False):
            self._host_driven = True
        self._step_rand = step_rand
        self._exec_mode = {"auto": _capi.B2_EXEC_AUTO, "persistent": _capi.B2_EXEC_PERSISTENT,
                           "lockstep": _capi.B2_EXEC_LOCKSTEP}[exec_mode]
        self._glm_path = {"auto": _capi.B2_GLM_AUTO, "group": _capi.B2_GLM_GROUP, "simt": _capi.B2_GLM_SIMT,
                          "tcgen05": _capi.B2_GLM_TCGEN05}[glm_path]
        self._warnings = []
        self._samples_after_tune = 0
        self._num_divs_sample = 0
        self._engine = None
        self._last_q = None

    # -- reference attribute surface
    @property
    def _logp_dlogp_func(self):
        if self.__logp_dlogp_func is None:
            self.__logp_dlogp_func = self._model.logp_dlogp_function(dtype="float64", device=self.device)
        return self.__logp_dlogp_func

    @property
    def vars_shape_dtype(self):
        return {vm.var: (vm.shp, vm.dtyp) for vm in self._ordering.vmap}

    def stop_tuning(self):
        self.tune = False

    def reset(self, start=None):
        self.tune = True
        self.potential.reset()

    # -- device options shared by the batched and the single-draw path
    def _opts(self):
        gamma, k, t0 = self._da
        o = dict(max_treedepth=getattr(self, "max_treedepth", 10),
                 early_max_treedepth=getattr(self, "early_max_treedepth", 8),
                 Emax=float(self.Emax), target_accept=float(self.target_accept), gamma=float(gamma),
                 k=float(k), t0=float(t0), adapt_step_size=int(bool(self.adapt_step_size)),
                 adapt_mass=int(self.potential.device_init()["adapt"]),
                 path_length=float(getattr(self, "path_length", 2.0)), max_steps=int(getattr(self, "max_steps", 1024)),
                 hmc_jitter=int(self._step_rand is not None), exec_mode=self._exec_mode, glm_path=self._glm_path)
        return o

    def _make_engine(self, n_chains, device=None):
        return self._model.engine(n_chains, dtype=self.dtype.name, device=self.device if device is None else device)

    def _init_engine_state(self, engine, q0, seeds):
        init = self.potential.device_init()
        if self._dense:
            engine.set_dense_mass(self.potential.device_chol())
        engine.set_state(q0, seeds, self._initial_step, init["mean"], init["var"], init["weight"], init["window"])

    # -- one chain, one draw (arraystep.py:258-264 + base_hmc.py:133-199)
    def step(self, point):
        q0 = self._model.dict_to_array(point)
        q, stats = self.astep(q0)
        return self._model.array_to_dict(q), stats

    @property
    def _batched(self):
        """can sampling.sample() run all chains of this step method in one device engine?"""
        return not self._host_driven

    def _host_astep(self, q0):
        """base_hmc.py:133-199 for a user potential: one transition on the host (host_transition.py)."""
        from . import host_transition as ht
        integ = ht.HostIntegrator(self.potential, lambda q: self._logp_dlogp_func(np.asarray(q, dtype="f8")))
        p0 = np.asarray(self.potential.random(), dtype="f8")
        start = integ.start(q0, p0)
        if not np.isfinite(start.energy):
            self.potential.raise_ok(self._ordering.vmap)
            self._warnings.append(SamplerWarning(WarningType.BAD_ENERGY, "Bad initial energy, check any log probabilities "
                                                 "that are inf or -inf, nan or very small", "critical", self.iter_count, None, None))
            raise SamplingError("Bad initial energy")
        adapt = self.tune and self.adapt_step_size
        step_size = float(self.step_adapt.current(adapt))
        self.step_size = step_size
        if self._step_rand is not None:
            step_size = float(self._step_rand(step_size))
        if self._kind == _capi.B2_NUTS:
            depth = self.early_max_treedepth if (self.tune and self.iter_count < 200) else self.max_treedepth
            q, grad, stats = ht.nuts_transition(integ, start, step_size, depth, float(self.Emax))
            accept = stats["mean_tree_accept"]
            if not self.tune and stats["depth"] >= depth and not stats["diverging"]:
                self._reached_max_treedepth += 1
        else:
            q, grad, stats = ht.hmc_transition(integ, start, step_size, float(self.path_length), int(self.max_steps), float(self.Emax))
            accept = stats["accept"]
        self.step_adapt.update(accept, adapt)
        self.potential.update(q, grad, self.tune)
        stats["tune"] = bool(self.tune)
        stats.update(self.step_adapt.stats())
        it = self.iter_count
        row = {k: np.dtype(dt).type(stats[k]) for k, dt in self.stats_dtypes[0].items()}
        self._account(row, it)
        return np.asarray(q, dtype="f8"), [row]

    def astep(self, q0):
        q0 = np.asarray(q0, dtype="f8")
        if self._host_driven:
            return self._host_astep(q0)
        if self._engine is None:
            self._engine = self._make_engine(1)
            seed = np.random.randint(2 ** 30)          # the reference draws from the global stream
            self._init_engine_state(self._engine, q0.reshape(1, -1), [seed])
        elif self._last_q is None or not np.array_equal(q0.astype(self.dtype), self._last_q):
            self._engine.set_position(q0.reshape(1, -1))
        it = self._engine.iter_done
        out = self._engine.run(self._kind, 1, it + 1 if self.tune else 0, self._opts())
        rep = self._engine.reports()[0]
        if rep.phase == _capi.PHASE_FAILED:
            self._raise_bad_energy(0, self._engine.mass_var()[0])
        host = {k: v.cpu().numpy()[0, 0] for k, v in out.items() if k != "q"}
        q = out["q"].cpu().numpy()[0, 0]
        self._last_q = q.astype(self.dtype)
        stats = self._stats_row(host)
        self._account(stats, it)
        self.step_size = float(host["step_size"])
        return q.astype("f8"), [stats]

    def _stats_row(self, host):
        row = {}
        for key, dt in self.stats_dtypes[0].items():
            if key == "path_length":
                row[key] = float(self.path_length)
            else:
                row[key] = np.dtype(dt).type(host[key])
        return row

    def _account(self, stats, it):
        if bool(stats["diverging"]):
            kind = WarningType.TUNING_DIVERGENCE if self.tune else WarningType.DIVERGENCE
            if not self.tune:
                self._num_divs_sample += 1
            self._warnings.append(SamplerWarning(kind, "Energy change in leapfrog step is too large.", "debug",
                                                 it, None, None))
        self.iter_count += 1
        if not self.tune:
            self._samples_after_tune += 1
            acc = stats.get("mean_tree_accept", stats.get("accept"))
            self.step_adapt._tuned_stats.append(float(acc))

    def _raise_bad_energy(self, chain, mass_var):
        """base_hmc.py:138-158."""
        if hasattr(self.potential, "sync"):
            self.potential.sync(mass_var)
            self.potential.raise_ok(self._ordering.vmap)
        msg = ("Bad initial energy, check any log probabilities that are inf or -inf, nan or very small "
               "(chain %d)" % chain)
        self._warnings.append(SamplerWarning(WarningType.BAD_ENERGY, msg, "critical", self.iter_count, None, None))
        raise SamplingError("Bad initial energy")

    # -- warnings (base_hmc.py:205-235)
    def _divergence_summary(self, n_divs, n_post):
        message = ""
        if n_divs and n_post == n_divs:
            message = "The chain contains only diverging samples. The model is probably misspecified."
        elif n_divs == 1:
            message = "There was 1 divergence after tuning. Increase `target_accept` or reparameterize."
        elif n_divs > 1:
            message = ("There were %s divergences after tuning. Increase `target_accept` or reparameterize."
                       % n_divs)
        if message:
            return [SamplerWarning(WarningType.DIVERGENCES, message, "error", None, None, None)]
        return []

    def warnings(self):
        warnings = self._warnings[:]
        warnings.extend(self._divergence_summary(self._num_divs_sample, self._samples_after_tune))
        warnings.extend(self.step_adapt.warnings())
        return warnings

    def _chain_warnings(self, report, mean_accept_post, n_post, diverging_rows, tune_flags, accept_ok=None):
        """Warnings of one chain of a batched run, from its device report and bulk reductions of its stats
        (`mean_accept_post`: mean acceptance statistic of the `n_post` post-tuning draws; `diverging_rows`: indices)."""
        warns = []
        for i in diverging_rows:
            kind = WarningType.TUNING_DIVERGENCE if tune_flags[i] else WarningType.DIVERGENCE
            warns.append(SamplerWarning(kind, "Energy change in leapfrog step is too large.", "debug",
                                        int(i), None, None))
        warns.extend(self._divergence_summary(report.n_div_post, report.n_post))
        warns.extend(step_sizes.acceptance_warnings(mean_accept_post, n_post, self.target_accept, ok=accept_ok))
        return warns
