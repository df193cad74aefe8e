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
