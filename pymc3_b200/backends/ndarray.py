"""In-memory per-chain trace.  API of pymc3/backends/ndarray.py:180-384 (NDArray) and the
save_trace/load_trace directory format of :32-177 (metadata.json + samples.npz per chain).

Difference in mechanism: the reference fills one row per draw by calling a Theano function on
the point (`record`, :258-277).  Here the engine accumulates the whole chain-batched trace on
the device; `NDArray.from_arrays` adopts *views* of the bulk host arrays, and back-transformed
deterministics are computed vectorised (Model.expand).  `setup/record/close` are kept for the
one-draw-at-a-time `step.step(point)` loop (sampling.iter_sample).
"""
import json
import os
import shutil

import numpy as np

from . import base


class NDArray(base.BaseTrace):
    supports_sampler_stats = True

    def __init__(self, name=None, model=None, vars=None, test_point=None):
        super().__init__(name, model, vars, test_point)
        self.draw_idx = 0
        self.draws = None
        self.samples = {}
        self._stats = None

    # ------------------------------------------------------------- bulk (device) path
    @classmethod
    def from_arrays(cls, model, chain, samples, stats):
        """samples: {varname: [draws, *shape]}, stats: {stat: [draws]} for one chain."""
        self = cls()                          # no model introspection: names, shapes and dtypes come from the arrays
        self.model = model                    # (thousands of chains are adopted per run)
        self.chain = chain
        self.samples = dict(samples)
        self.varnames = list(samples.keys())
        self.var_shapes = {k: v.shape[1:] for k, v in samples.items()}
        self.var_dtypes = {k: v.dtype for k, v in samples.items()}
        n = len(next(iter(samples.values()))) if samples else 0
        self.draws = self.draw_idx = n
        self._stats = [dict(stats)] if stats is not None else None
        self.sampler_vars = [{k: v.dtype for k, v in stats.items()}] if stats is not None else None
        return self

    # ------------------------------------------------------------- per-draw path
    def setup(self, draws, chain, sampler_vars=None):
        super().setup(draws, chain, sampler_vars)
        self.chain = chain
        if self.samples:                      # extend an existing trace (ndarray.py:221-231)
            old = self.draws
            self.draws = old + draws
            self.draw_idx = old
            for name, val in self.samples.items():
                self.samples[name] = np.concatenate([val, np.zeros((draws,) + val.shape[1:], val.dtype)])
        else:
            self.draws = draws
            for name in self.varnames:
                self.samples[name] = np.zeros((draws,) + tuple(self.var_shapes[name]), dtype=self.var_dtypes[name])
        if sampler_vars is None:
            return
        if self._stats is None:
            self._stats = [{k: np.zeros(draws, dtype=dt) for k, dt in sampler.items()} for sampler in sampler_vars]
        else:
            for data, sampler in zip(self._stats, sampler_vars):
                if set(sampler) != set(data):
                    raise ValueError("Sampler vars can't change")
                for k in data:
                    data[k] = np.concatenate([data[k], np.zeros(draws, dtype=data[k].dtype)])

    def record(self, point, sampler_stats=None):
        full = self.model.expand(self.model.dict_to_array(point)) if self.model is not None else point
        for name in self.varnames:
            self.samples[name][self.draw_idx] = full[name]
        if self._stats is not None and sampler_stats is None:
            raise ValueError("Expected sampler_stats")
        if self._stats is None and sampler_stats is not None:
            raise ValueError("Unknown sampler_stats")
        if sampler_stats is not None:
            for data, vars in zip(self._stats, sampler_stats):
                for key, val in vars.items():
                    data[key][self.draw_idx] = val
        self.draw_idx += 1

    def _get_sampler_stats(self, varname, sampler_idx, burn, thin):
        return self._stats[sampler_idx][varname][burn::thin]

    def close(self):
        if self.draw_idx == self.draws:
            return
        self.samples = {name: val[: self.draw_idx] for name, val in self.samples.items()}
        if self._stats is not None:
            self._stats = [{k: v[: self.draw_idx] for k, v in st.items()} for st in self._stats]
        self.draws = self.draw_idx

    # ------------------------------------------------------------- selection
    def __len__(self):
        if not self.samples:
            return 0
        return self.draw_idx

    def get_values(self, varname, burn=0, thin=1):
        return self.samples[varname][burn::thin]

    def _slice(self, idx):
        if idx.start in (None, 0) and idx.stop is None and idx.step in (None, 1):
            return self
        sliced = NDArray()
        sliced.model = self.model
        sliced.chain = self.chain
        sliced.varnames = list(self.varnames)
        sliced.var_shapes, sliced.var_dtypes = self.var_shapes, self.var_dtypes
        sliced.samples = {name: val[idx] for name, val in self.samples.items()}
        sliced.sampler_vars = self.sampler_vars
        sliced.draws = sliced.draw_idx = len(next(iter(sliced.samples.values()))) if sliced.samples else 0
        if self._stats is not None:
            sliced._stats = [{k: v[idx] for k, v in st.items()} for st in self._stats]
        return sliced

    def point(self, idx):
        idx = int(idx)
        return {name: values[idx] for name, values in self.samples.items()}


# ---------------------------------------------------------------- on-disk format (N3 row)
def save_trace(trace, directory=None, overwrite=False):
    """One sub-directory per chain with metadata.json + samples.npz (ndarray.py:32-73)."""
    if directory is None:
        directory = ".pymc_{}.trace"
        idx = 1
        while os.path.exists(directory.format(idx)):
            idx += 1
        directory = directory.format(idx)
    if os.path.isdir(directory):
        if overwrite:
            shutil.rmtree(directory)
        else:
            raise OSError("Cautiously refusing to overwrite the already existing {}! Please supply "
                          "a different directory, or set `overwrite=True`".format(directory))
    os.makedirs(directory)
    for chain, ndarray in trace._straces.items():
        cdir = os.path.join(directory, str(chain))
        os.mkdir(cdir)
        stats = None
        if ndarray._stats is not None:
            stats = [{k: np.asarray(v).tolist() for k, v in st.items()} for st in ndarray._stats]
        meta = {"draw_idx": int(ndarray.draw_idx), "draws": int(ndarray.draws), "_stats": stats,
                "chain": int(ndarray.chain)}
        with open(os.path.join(cdir, "metadata.json"), "w") as f:
            json.dump(meta, f)
        np.savez(os.path.join(cdir, "samples.npz"), **ndarray.samples)
    return directory


def load_trace(directory, model=None):
    """ndarray.py:75-95."""
    straces = []
    for sub in sorted(os.listdir(directory), key=lambda s: int(s) if s.isdigit() else -1):
        cdir = os.path.join(directory, sub)
        if not os.path.isdir(cdir):
            continue
        with open(os.path.join(cdir, "metadata.json")) as f:
            meta = json.load(f)
        with np.load(os.path.join(cdir, "samples.npz")) as z:
            samples = {k: z[k] for k in z.files}
        stats = None
        if meta["_stats"] is not None:
            stats = {k: np.array(v) for k, v in meta["_stats"][0].items()}
        st = NDArray.from_arrays(model, meta["chain"], samples, stats)
        straces.append(st)
    return base.MultiTrace(straces)
