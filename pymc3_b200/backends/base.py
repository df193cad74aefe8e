"""Trace containers with the reference's selection API (pymc3/backends/base.py:39-559):
`MultiTrace.get_values(varname, burn, thin, combine, chains, squeeze)`, `get_sampler_stats`,
`trace['x']`, `trace['x', 100::2]`, `trace[100:]`, `trace.point(i, chain)`, `points()`,
`varnames`, `stat_names`, `chains`, `nchains`, `report`; and `merge_traces` (:562-584).

The containers hold host arrays only; the engine writes draws on the device and hands over
bulk arrays once per run (see ndarray.NDArray.from_arrays).
"""
import itertools
import logging

import numpy as np

from .report import SamplerReport, merge_reports

logger = logging.getLogger("pymc3")


class BackendError(Exception):
    pass


def _check_stat_dtypes(sampler_vars):
    """A statistic reported by several samplers must have one dtype (base.py:113-128)."""
    seen = {}
    for per_sampler in sampler_vars:
        for stat, dtype in per_sampler.items():
            if seen.setdefault(stat, dtype) != dtype:
                raise ValueError("Sampler statistic %s appears with different types." % stat)


class BaseTrace:
    """Draws of one chain; storage is up to the sub-class."""

    supports_sampler_stats = False

    def __init__(self, name=None, model=None, vars=None, test_point=None):
        self.name, self.model, self.chain = name, model, None
        self.sampler_vars = None
        self._is_base_setup = False
        self._warnings = []
        self.varnames, self.var_shapes, self.var_dtypes = [], {}, {}
        if model is not None:
            # every unobserved RV is traced: free (transformed) variables and deterministics
            example = model.expand(model.dict_to_array(model.test_point))
            self.varnames = list(vars) if vars is not None else list(model.unobserved_RVs)
            for n in self.varnames:
                arr = np.asarray(example[n])
                self.var_shapes[n], self.var_dtypes[n] = arr.shape, arr.dtype

    # -- life cycle used by the draw-at-a-time loop
    def setup(self, draws, chain, sampler_vars=None):
        if sampler_vars is not None:
            if not self.supports_sampler_stats:
                raise ValueError("Backend does not support sampler stats.")
            _check_stat_dtypes(sampler_vars)
        if self._is_base_setup and self.sampler_vars != sampler_vars:
            raise ValueError("Can't change sampler_vars")
        self.sampler_vars = sampler_vars
        self._is_base_setup = True

    def record(self, point, sampler_states=None):
        raise NotImplementedError

    def close(self):
        pass

    def _add_warnings(self, warnings):
        self._warnings.extend(warnings)

    # -- selection
    def __len__(self):
        raise NotImplementedError

    def __getitem__(self, key):
        if isinstance(key, slice):
            return self._slice(key)
        try:
            return self.point(int(key))
        except (ValueError, TypeError):
            raise ValueError("Can only index with slice or integer")

    def get_values(self, varname, burn=0, thin=1):
        raise NotImplementedError

    def get_sampler_stats(self, stat_name, sampler_idx=None, burn=0, thin=1):
        """One array per draw; if several samplers report the stat, they are stacked last."""
        if not self.supports_sampler_stats:
            raise ValueError("This backend does not support sampler stats")
        if sampler_idx is not None:
            return self._get_sampler_stats(stat_name, sampler_idx, burn, thin)
        owners = [i for i, per_sampler in enumerate(self.sampler_vars or []) if stat_name in per_sampler]
        if not owners:
            raise KeyError("Unknown sampler stat %s" % stat_name)
        cols = [self._get_sampler_stats(stat_name, i, burn, thin) for i in owners]
        return cols[0] if len(cols) == 1 else np.stack(cols, axis=-1)

    @property
    def stat_names(self):
        out = set()
        for per_sampler in (self.sampler_vars or []):
            out.update(per_sampler)
        return out


class MultiTrace:
    """The chains of one run, keyed by chain id."""

    def __init__(self, straces):
        self._straces = {}
        self._report = SamplerReport()
        for st in straces:
            if st.chain in self._straces:
                raise ValueError("Chains are not unique.")
            self._straces[st.chain] = st
            self._report._add_warnings(getattr(st, "_warnings", []), st.chain)

    def __repr__(self):
        return "<MultiTrace: %d chains, %d iterations, %d variables>" % (self.nchains, len(self), len(self.varnames))

    # -- shape
    @property
    def chains(self):
        return sorted(self._straces)

    @property
    def nchains(self):
        return len(self._straces)

    @property
    def report(self):
        return self._report

    def _last(self):
        return self._straces[self.chains[-1]]

    def __len__(self):
        return len(self._last())

    @property
    def varnames(self):
        return self._last().varnames

    @property
    def stat_names(self):
        per_chain = [st.sampler_vars for st in self._straces.values()]
        if any(sv != per_chain[0] for sv in per_chain):
            raise ValueError("Inividual chains contain different sampler stats")
        out = set()
        for st in self._straces.values():
            out |= st.stat_names
        return out

    # -- selection
    def __getitem__(self, key):
        if isinstance(key, slice):
            return self._slice(key)
        if isinstance(key, (int, np.integer)):
            return self.point(int(key))
        burn, thin = 0, 1
        if isinstance(key, tuple):
            key, sl = key
            burn, thin = sl.start or 0, sl.step or 1
        key = str(key)
        in_vars, in_stats = key in self.varnames, key in self.stat_names
        if in_vars and in_stats:
            logger.warning("Attribute access on a trace object is ambigous. Sampler statistic and model "
                           "variable share a name. Use trace.get_values or trace.get_sampler_stats.")
        if in_vars:
            return self.get_values(key, burn=burn, thin=thin)
        if in_stats:
            return self.get_sampler_stats(key, burn=burn, thin=thin)
        raise KeyError("Unknown variable %s" % key)

    def __getattr__(self, name):
        if name.startswith("_") or name in ("varnames", "chains", "stat_names"):
            raise AttributeError(name)
        if name in self.varnames:
            return self.get_values(name)
        if name in self.stat_names:
            return self.get_sampler_stats(name)
        raise AttributeError("'%s' object has no attribute '%s'" % (type(self).__name__, name))

    def _chain_list(self, chains):
        if chains is None:
            return self.chains
        return list(chains) if np.iterable(chains) else [chains]

    def get_values(self, varname, burn=0, thin=1, combine=True, chains=None, squeeze=True):
        parts = [self._straces[c].get_values(str(varname), burn, thin) for c in self._chain_list(chains)]
        return _squeeze_cat(parts, combine, squeeze)

    def get_sampler_stats(self, stat_name, burn=0, thin=1, combine=True, chains=None, squeeze=True):
        if stat_name == "tree_depth" and "depth" in self.stat_names:
            stat_name = "depth"                # BASELINE.json's name for the reference's `depth` stat
        if stat_name not in self.stat_names:
            raise KeyError("Unknown sampler statistic %s" % stat_name)
        parts = [self._straces[c].get_sampler_stats(stat_name, None, burn, thin) for c in self._chain_list(chains)]
        return _squeeze_cat(parts, combine, squeeze)

    def _slice(self, sl):
        out = MultiTrace([st._slice(sl) for st in self._straces.values()])
        out._report = self._report._slice(*sl.indices(len(self)))
        return out

    def point(self, idx, chain=None):
        return self._straces[self.chains[-1] if chain is None else chain].point(idx)

    def points(self, chains=None):
        return itertools.chain.from_iterable(self._straces[c] for c in self._chain_list(chains))


def merge_traces(mtraces):
    """Fold the chains of later MultiTraces into the first; chain ids must be unique."""
    first = mtraces[0]
    for other in mtraces[1:]:
        for cid, st in other._straces.items():
            if cid in first._straces:
                raise ValueError("Chains are not unique.")
            first._straces[cid] = st
    first._report = merge_reports([mt.report for mt in mtraces])
    return first


def _squeeze_cat(results, combine, squeeze):
    """combine -> one concatenated array (wrapped in a list if not squeeze);
    else a list per chain, unwrapped when it has one element and squeeze is set."""
    if combine:
        joined = np.concatenate(results)
        return joined if squeeze else [joined]
    if squeeze and len(results) == 1:
        return results[0]
    return results
