"""Sampler warnings and the per-trace report.  Mirrors pymc3/backends/report.py:25-217:
WarningType / SamplerWarning, convergence checks (R-hat > 1.05/1.2/1.4, ESS < 200 / 10% / 25%)
with ESS / R-hat from pymc3_b200.stats (arviz is not available), and the same log summary."""
import collections
import enum
import logging

import numpy as np

logger = logging.getLogger("pymc3")


class WarningType(enum.Enum):
    DIVERGENCE = 1
    TUNING_DIVERGENCE = 2
    DIVERGENCES = 3
    TREEDEPTH = 4
    BAD_PARAMS = 5
    CONVERGENCE = 6
    BAD_ACCEPTANCE = 7
    BAD_ENERGY = 8


SamplerWarning = collections.namedtuple("SamplerWarning", "kind, message, level, step, exec_info, extra")
_LEVELS = {"info": logging.INFO, "error": logging.ERROR, "warn": logging.WARN, "debug": logging.DEBUG,
           "critical": logging.CRITICAL}


class SamplerReport:
    def __init__(self):
        self._chain_warnings = {}
        self._global_warnings = []
        self._ess = None
        self._rhat = None
        self._n_tune = None
        self._n_draws = None
        self._t_sampling = None

    @property
    def _warnings(self):
        chains = sum(self._chain_warnings.values(), [])
        return chains + self._global_warnings

    @property
    def ok(self):
        """Whether the automatic convergence checks found serious problems."""
        return all(_LEVELS[warn.level] < _LEVELS["warn"] for warn in self._warnings)

    @property
    def n_tune(self):
        return self._n_tune

    @property
    def n_draws(self):
        return self._n_draws

    @property
    def t_sampling(self):
        return self._t_sampling

    def raise_ok(self, level="error"):
        errors = [warn for warn in self._warnings if _LEVELS[warn.level] >= _LEVELS[level]]
        if errors:
            raise ValueError("Serious convergence issues during sampling.")

    # thresholds of pymc3/backends/report.py:126-166, table-driven
    _RHAT_RULES = ((1.4, "error", "The rhat statistic is larger than 1.4 for some parameters. "
                                  "The sampler did not converge."),
                   (1.2, "warn", "The rhat statistic is larger than 1.2 for some parameters."),
                   (1.05, "info", "The rhat statistic is larger than 1.05 for some parameters. This "
                                  "indicates slight problems during sampling."))

    def _run_convergence_checks(self, trace, model):
        from .. import stats
        if trace.nchains == 1:
            note = "Only one chain was sampled, this makes it impossible to run some convergence checks"
            self._add_warnings([SamplerWarning(WarningType.BAD_PARAMS, note, "info", None, None, None)])
            return
        self._ess, self._rhat = {}, {}
        for name in trace.varnames:
            if name.endswith("__"):            # transformed twins are reported on the natural scale
                continue
            draws = np.stack(trace.get_values(name, combine=False))
            self._ess[name] = stats.ess(draws)
            self._rhat[name] = stats.rhat(draws)
        found = []
        worst_rhat = max(float(np.max(v)) for v in self._rhat.values())
        for bound, level, text in self._RHAT_RULES:
            if worst_rhat > bound:
                found.append(SamplerWarning(WarningType.CONVERGENCE, text, level, None, None, self._rhat))
                break
        worst_ess = min(float(np.min(v)) for v in self._ess.values())
        total = len(trace) * trace.nchains
        ess_msg = None
        if worst_ess < 200 and total >= 500:
            ess_msg = ("error", "The estimated number of effective samples is smaller than 200 for some "
                                "parameters.")
        elif worst_ess / total < 0.1:
            ess_msg = ("warn", "The number of effective samples is smaller than 10% for some parameters.")
        elif worst_ess / total < 0.25:
            ess_msg = ("info", "The number of effective samples is smaller than 25% for some parameters.")
        if ess_msg:
            found.append(SamplerWarning(WarningType.CONVERGENCE, ess_msg[1], ess_msg[0], None, None, self._ess))
        self._add_warnings(found)

    def _add_warnings(self, warnings, chain=None):
        target = self._global_warnings if chain is None else self._chain_warnings.setdefault(chain, [])
        target.extend(warnings)

    def _log_summary(self):
        """report.py:205-207 logs every warning; with thousands of chains the same kind of warning repeats per
        chain, so beyond 8 chains each kind is logged once with the number of chains it concerns."""
        if len(self._chain_warnings) <= 8:
            for warn in self._warnings:
                logger.log(_LEVELS[warn.level], warn.message)
            return
        by_kind = {}
        for chain, warns in self._chain_warnings.items():
            for warn in warns:
                if _LEVELS[warn.level] <= logging.DEBUG:
                    continue
                entry = by_kind.setdefault((warn.kind, warn.level), [set(), warn.message])
                entry[0].add(chain)
        for (kind, level), (chains_hit, example) in by_kind.items():
            logger.log(_LEVELS[level], "%s  [%d of %d chains]" % (example, len(chains_hit), len(self._chain_warnings)))
        for warn in self._global_warnings:
            logger.log(_LEVELS[warn.level], warn.message)

    def _slice(self, start, stop, step):
        """Report for trace[start:stop:step]: per-draw warnings are re-indexed or dropped."""
        def keep(warn):
            if warn.step is None:
                return warn
            if start <= warn.step < stop and (warn.step - start) % step == 0:
                return warn._replace(step=warn.step - start)
            return None

        out = SamplerReport()
        out._add_warnings([w for w in map(keep, self._global_warnings) if w is not None])
        for chain, warns in self._chain_warnings.items():
            out._add_warnings([w for w in map(keep, warns) if w is not None], chain)
        return out


def merge_reports(reports):
    total = SamplerReport()
    for report in reports:
        for chain, warns in report._chain_warnings.items():
            total._add_warnings(warns, chain)
        total._add_warnings(report._global_warnings)
    return total
