"""Python handle on one device engine (one per GPU).  Plumbing only: PyTorch owns the device
buffers, ctypes calls libb200nuts.so; no arithmetic of the hot path happens here.

Mirrors the runtime half of the reference's model->array bridge
(pymc3/model.py:541-713 ValueGradFunction) plus the draw loop's state
(pymc3/sampling.py:847-936), batched over chains.
"""
import ctypes as C

import numpy as np

from . import _capi

_NUTS_STATS = {"energy": "f8", "energy_error": "f8", "max_energy_error": "f8", "mean_tree_accept": "f8",
               "step_size": "f8", "step_size_bar": "f8", "model_logp": "f8", "depth": "i4",
               "tree_size": "i4", "diverging": "u1", "tune": "u1"}
_HMC_STATS = {"energy": "f8", "energy_error": "f8", "step_size": "f8", "step_size_bar": "f8",
              "model_logp": "f8", "accept": "f8", "n_steps": "i4", "diverging": "u1", "tune": "u1",
              "accepted": "u1"}


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise _capi.B2Error("pymc3_b200 needs a CUDA device: the NUTS/HMC hot path has no CPU fallback")
    return torch


class Engine:
    """`n_chains` chains of one model family on one CUDA device."""

    def __init__(self, desc_builder, n_chains, dtype="float32", device=0):
        torch = _torch()
        self.torch = torch
        self.lib = _capi.load_library()
        self.device = int(device)
        self.dev = torch.device("cuda", self.device)
        self.n_chains = int(n_chains)
        self.np_dtype = np.dtype(dtype)
        if self.np_dtype not in (np.dtype("float32"), np.dtype("float64")):
            raise TypeError("Invalid dtype %s: engine vectors are float32 or float64" % dtype)
        self.t_dtype = torch.float64 if self.np_dtype == np.float64 else torch.float32
        self._keep = []                       # device tensors referenced by raw pointer
        self.desc = desc_builder(self._upload)
        self.D = int(self.desc.D)
        handle = C.c_void_p()
        _capi.check(self.lib.b2_engine_create(C.byref(self.desc), self.n_chains,
                                              _capi.B2_F64 if self.np_dtype == np.float64 else _capi.B2_F32,
                                              self.device, C.byref(handle)), self.lib)
        self.handle = handle
        self.iter_done = 0
        self._chol = self._chol_t = None

    # -- plumbing
    def _upload(self, array, dtype):
        torch = self.torch
        if isinstance(array, torch.Tensor):            # data generated on the device stays there
            tdt = {"f4": torch.float32, "f8": torch.float64, np.uint8: torch.uint8, np.int32: torch.int32}[dtype]
            t = array.to(device=self.dev, dtype=tdt).contiguous()
        else:
            t = torch.as_tensor(np.ascontiguousarray(array, dtype=dtype), device=self.dev)
        self._keep.append(t)
        return t.data_ptr()

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.dev).cuda_stream)

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle.value:
            self.lib.b2_engine_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- ValueGradFunction.__call__ for a batch of points
    def logp_dlogp(self, q, glm_path=_capi.B2_GLM_AUTO):
        """q: [n, D] host array or device tensor -> (logp [n] float64, grad [n, D]) device tensors."""
        torch = self.torch
        qd = torch.as_tensor(q, device=self.dev).to(self.t_dtype).contiguous()
        if qd.ndim != 2 or qd.shape[1] != self.D:
            raise ValueError("Invalid shape for array. Must be (n, %d) but is %s." % (self.D, tuple(qd.shape)))
        n = qd.shape[0]
        if n > self.n_chains:
            raise ValueError("at most n_chains=%d points per call" % self.n_chains)
        logp = torch.empty(n, dtype=torch.float64, device=self.dev)
        grad = torch.empty_like(qd)
        _capi.check(self.lib.b2_logp_dlogp(self.handle, qd.data_ptr(), n, logp.data_ptr(), grad.data_ptr(),
                                           int(glm_path), self._stream()), self.lib)
        return logp, grad

    # -- CpuLeapfrogIntegrator.compute_state + n x .step (integration.py:39-109), all chains at once
    def leapfrog(self, q, p, var, epsilon, n_steps, glm_path=_capi.B2_GLM_AUTO):
        """q, p: [n_chains, D]; var: [D] diagonal of M^-1 -> (q', p', energy [n_chains]) device tensors after
        `n_steps` leapfrog steps of size `epsilon` (negative = backwards).  Overwrites the sampler state."""
        torch = self.torch
        qd = torch.as_tensor(q, device=self.dev).to(self.t_dtype).contiguous()
        pd = torch.as_tensor(p, device=self.dev).to(self.t_dtype).contiguous()
        if qd.shape != (self.n_chains, self.D) or pd.shape != qd.shape:
            raise ValueError("q and p must have shape (%d, %d)" % (self.n_chains, self.D))
        vd = torch.as_tensor(np.ascontiguousarray(var, dtype="f8"), device=self.dev)
        if vd.numel() != self.D:
            raise ValueError("var must have %d elements" % self.D)
        q_out, p_out = torch.empty_like(qd), torch.empty_like(pd)
        energy = torch.empty(self.n_chains, dtype=torch.float64, device=self.dev)
        _capi.check(self.lib.b2_leapfrog(self.handle, qd.data_ptr(), pd.data_ptr(), vd.data_ptr(), float(epsilon),
                                         int(n_steps), q_out.data_ptr(), p_out.data_ptr(), energy.data_ptr(),
                                         int(glm_path), self._stream()), self.lib)
        return q_out, p_out, energy

    # -- QuadPotentialFull / FullInv (quadpotential.py:400-479) for every chain: the device integrates z = L^-1 q
    def set_dense_mass(self, chol):
        """chol: [D, D] lower Cholesky factor of the covariance (None switches the dense metric off).  The engine
        then samples z = L^-1 q with unit mass (b2_set_dense_mass); this class converts at its boundary, so callers
        keep handing over and receiving q: set_state / set_position take q, run() and position() return q."""
        torch = self.torch
        if chol is None:
            _capi.check(self.lib.b2_set_dense_mass(self.handle, None, self._stream()), self.lib)
            self._chol = self._chol_t = None
            return
        L = np.tril(np.asarray(chol, dtype="f8"))
        if L.shape != (self.D, self.D):
            raise ValueError("chol must have shape (%d, %d)" % (self.D, self.D))
        self._chol = L
        self._chol_t = torch.as_tensor(np.ascontiguousarray(L.T), device=self.dev).to(self.t_dtype)     # q = z @ L^T
        t = torch.as_tensor(np.ascontiguousarray(L), device=self.dev).to(self.t_dtype).contiguous()
        _capi.check(self.lib.b2_set_dense_mass(self.handle, t.data_ptr(), self._stream()), self.lib)
        torch.cuda.synchronize(self.dev)               # `t` is copied by the call; it may go now

    def _to_z(self, q):
        import scipy.linalg
        return scipy.linalg.solve_triangular(self._chol, np.asarray(q, dtype="f8").T, lower=True).T

    # -- sampling.py:410-413, 883-884, 1915-1929 + base_hmc.py:93-103
    def set_state(self, q0, seeds, step_size0, mass_mean, mass_var, mass_weight, adaptation_window=101):
        torch = self.torch
        if self._chol is not None:
            q0 = self._to_z(np.asarray(q0, dtype="f8").reshape(self.n_chains, self.D))
        q0 = np.asarray(q0, dtype=self.np_dtype).reshape(self.n_chains, self.D)
        self._q0 = torch.as_tensor(np.ascontiguousarray(q0), device=self.dev)
        self._seeds = torch.as_tensor(np.asarray(seeds, dtype=np.uint64).view(np.int64).copy(), device=self.dev)
        self._mm = torch.as_tensor(np.ascontiguousarray(mass_mean, dtype="f8"), device=self.dev)
        self._mv = torch.as_tensor(np.ascontiguousarray(mass_var, dtype="f8"), device=self.dev)
        if self._mm.numel() != self.D or self._mv.numel() != self.D:
            raise ValueError("mass mean/var must have %d elements" % self.D)
        _capi.check(self.lib.b2_set_state(self.handle, self._q0.data_ptr(), self._seeds.data_ptr(),
                                          float(step_size0), self._mm.data_ptr(), self._mv.data_ptr(),
                                          float(mass_weight), int(adaptation_window), self._stream()), self.lib)
        self.iter_done = 0

    def set_position(self, q):
        if self._chol is not None:
            q = self._to_z(np.asarray(q, dtype="f8").reshape(self.n_chains, self.D))
        q = np.asarray(q, dtype=self.np_dtype).reshape(self.n_chains, self.D)
        t = self.torch.as_tensor(np.ascontiguousarray(q), device=self.dev)
        _capi.check(self.lib.b2_set_position(self.handle, t.data_ptr(), self._stream()), self.lib)
        self.torch.cuda.synchronize(self.dev)

    # -- the draw loop for all chains (sampling.py:914-936)
    def alloc_trace(self, kind, n_iters, trace_q=True):
        """Device trace buffers for `n_iters` iterations: {'q': [n, C, D], stat: [n, C]}."""
        torch = self.torch
        names = _NUTS_STATS if kind == _capi.B2_NUTS else _HMC_STATS
        out = {}
        if trace_q:
            out["q"] = torch.empty((n_iters, self.n_chains, self.D), dtype=self.t_dtype, device=self.dev)
        for name, code in names.items():
            out[name] = torch.zeros((n_iters, self.n_chains), device=self.dev,
                                    dtype={"f8": torch.float64, "i4": torch.int32, "u1": torch.uint8}[code])
        return out

    def run(self, kind, n_iters, tune_until, opts, trace_q=True, out=None, row0=0, run_ahead=False):
        """Runs `n_iters` transitions of every chain; returns {'q': [n, C, D], stat: [n, C]} device
        tensors.  With `out` (from alloc_trace) the rows [row0, row0 + n_iters) of those buffers are
        filled instead of allocating (chunked runs of one long job).  `run_ahead` (needs `out`): in
        lock-step runs, chains that finish this chunk early go on into the following rows of `out`
        instead of idling until the slowest chain of the chunk is done; the call still returns as
        soon as every chain has completed the chunk, and the next chunk picks the chains up where
        they are."""
        if out is None:
            out = self.alloc_trace(kind, n_iters, trace_q)
            view = out
        else:
            view = {k: v[row0:row0 + n_iters] for k, v in out.items()}
        tr = _capi.TraceOut()
        for name, t in view.items():
            assert t.is_contiguous() and t.shape[0] == n_iters
            setattr(tr, "d_" + name, t.data_ptr())
        o = _capi.SamplerOpts(kind=kind, n_iters=int(n_iters), tune_until=int(tune_until), **opts)
        if run_ahead and out is not view:
            o.run_ahead = int(next(iter(out.values())).shape[0] - row0 - n_iters)
        _capi.check(self.lib.b2_sample_run(self.handle, C.byref(o), C.byref(tr), self._stream()), self.lib)
        self.iter_done += int(n_iters)
        if self._chol is not None and "q" in view:      # rows of this call are complete: z -> q = L z, in place
            view["q"].copy_(view["q"] @ self._chol_t)
        return view

    def reports(self):
        arr = (_capi.ChainReport * self.n_chains)()
        _capi.check(self.lib.b2_get_chain_reports(self.handle, arr), self.lib)
        return list(arr)

    def mass_var(self):
        out = np.empty((self.n_chains, self.D))
        _capi.check(self.lib.b2_get_mass_var(self.handle, out.ctypes.data), self.lib)
        return out

    def position(self):
        out = np.empty((self.n_chains, self.D))
        _capi.check(self.lib.b2_get_position(self.handle, out.ctypes.data), self.lib)
        return out @ self._chol.T if self._chol is not None else out

    def set_profiling(self, on=True):
        _capi.check(self.lib.b2_set_profiling(self.handle, int(bool(on))), self.lib)

    def profile(self):
        """(total ms, launches) of the chain-batched likelihood kernel since set_profiling(True)."""
        ms, n = C.c_double(), C.c_int64()
        _capi.check(self.lib.b2_get_profile(self.handle, C.byref(ms), C.byref(n)), self.lib)
        return ms.value, n.value

    def profile_advance(self):
        """total ms of the advance (state machine) kernel over the launches counted by profile()."""
        ms = C.c_double()
        _capi.check(self.lib.b2_get_profile_advance(self.handle, C.byref(ms)), self.lib)
        return ms.value

    def kernel_launches(self):
        return int(self.lib.b2_kernel_launches(self.handle))
