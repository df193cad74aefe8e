"""World-size-2 `gloo` test of the multi-GPU host logic (chain sharding is the N>1 path)."""
import os
import subprocess
import sys
import textwrap

import numpy as np

from pymc3_b200 import distributed as b2d

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shards_cover_all_chains_and_seeds_are_global():
    for total, world in [(1024, 8), (10, 4), (7, 2), (4096, 8)]:
        parts = [b2d.shard_chains(total, world, r) for r in range(world)]
        assert parts[0][0] == 0 and parts[-1][1] == total
        assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
        assert max(hi - lo for lo, hi in parts) - min(hi - lo for lo, hi in parts) <= 1
        seeds = np.concatenate([b2d.global_chain_seeds(99, lo, hi) for lo, hi in parts])
        assert np.array_equal(seeds, b2d.global_chain_seeds(99, 0, total))


def test_two_rank_gloo_reductions(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(textwrap.dedent("""
        import os, sys, json
        sys.path.insert(0, %r)
        import numpy as np
        import torch.distributed as dist
        from pymc3_b200 import distributed as b2d
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        lo, hi = b2d.shard_chains(10, world, rank)
        secs, counts = b2d.reduce_job_metrics([1.0 + rank, 5.0 - rank], [hi - lo, 100 * (rank + 1)])
        mn, vec = b2d.combine_ess([10.0 + rank, 3.0, 50.0])
        if rank == 0:
            print(json.dumps({"secs": secs, "counts": counts, "min": mn, "vec": vec.tolist()}))
        dist.destroy_process_group()
    """ % ROOT))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29631", str(script)],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(line)
    assert res["secs"] == [2.0, 5.0] and res["counts"] == [10.0, 300.0]
    assert res["min"] == 6.0 and res["vec"] == [21.0, 6.0, 100.0]


def test_sharded_lockstep_driver_over_gloo(tmp_path):
    """The observation-sharded driver loop (pymc3_b200.sharded.run_lockstep_sharded) with its collective
    over gloo, world size 2, against a stand-in engine: both ranks must apply identical reduced values and
    stop together."""
    script = tmp_path / "worker2.py"
    script.write_text(textwrap.dedent("""
        import sys, json, ctypes
        sys.path.insert(0, %r)
        import numpy as np, torch
        import torch.distributed as dist
        from pymc3_b200 import sharded, _capi
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()

        class FakeLib:                       # stands in for libb200nuts.so: partial sums of a row-sharded quadratic
            def __init__(self, eng): self.e = eng
            def b2_step_begin(self, *a): return 0
            def b2_step_likelihood(self, h, ptr, s):
                e = self.e
                e.packed[:] = torch.from_numpy(np.outer(np.arange(1, e.n_chains + 1), e.rows.sum(0))).double() * (e.it + 1)
                return 0
            def b2_step_advance(self, h, ptr, copies, s):
                e = self.e
                assert copies == world
                e.seen.append(e.packed.clone())
                e.it += 1
                return 0
            def b2_step_active(self, h, cnt, s):
                cnt._obj.value = 0 if self.e.it >= 32 else 1
                return 0
            def b2_step_end(self, h): return 0
            def b2_last_error(self): return b""

        class FakeEngine:
            def __init__(self):
                lo, hi = sharded.row_shard(10, world, rank)
                self.rows = np.arange(40, dtype="f8").reshape(10, 4)[lo:hi]
                self.n_chains, self.D, self.it, self.seen = 3, 3, 0, []
                self.torch, self.dev, self.handle, self.iter_done = torch, "cpu", None, 0
                self.lib = FakeLib(self)
            def _stream(self): return None
            def alloc_trace(self, kind, n): return {}

        eng = FakeEngine()
        orig_zeros = torch.zeros
        def zeros(*a, **k):
            t = orig_zeros(*a, **k); eng.packed = t; return t
        torch.zeros = zeros
        sharded.run_lockstep_sharded(eng, 0, 5, 5, dict(max_treedepth=10, early_max_treedepth=8, Emax=1000.0,
            target_accept=0.8, gamma=0.05, k=0.75, t0=10.0, adapt_step_size=1, adapt_mass=1, path_length=2.0,
            max_steps=1024, hmc_jitter=0, exec_mode=2, glm_path=0), lambda t: dist.all_reduce(t), world)
        full = np.outer(np.arange(1, 4), np.arange(40, dtype="f8").reshape(10, 4).sum(0))
        ok = all(np.allclose(s.numpy(), full * (i + 1)) for i, s in enumerate(eng.seen))
        gathered = [None] * world
        dist.all_gather_object(gathered, (len(eng.seen), ok, eng.last_lockstep_steps))
        if rank == 0:
            print(json.dumps(gathered))
        dist.destroy_process_group()
    """ % ROOT))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29641", str(script)],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-3000:]
    import json
    line = [l for l in out.stdout.splitlines() if l.startswith("[")][-1]
    res = json.loads(line)
    assert res == [[32, True, 32], [32, True, 32]]
