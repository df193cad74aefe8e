"""Multi-GPU plumbing: one process per GPU, chains sharded with NO data-path collective.

The reference's only parallelism is one OS process per chain (pymc3/parallel_sampling.py:353-445)
with no inter-chain communication; chains are therefore the shard unit.  Seeds and
start points are keyed by the GLOBAL chain id, so results do not depend on the GPU count.  The only
collectives are the bookkeeping reductions below (timings as max over ranks, counts as sums); they
run over whatever backend the process group was created with (`nccl` on GPUs, `gloo` in CPU tests).
"""
import numpy as np


def shard_chains(total_chains, world_size, rank):
    """Contiguous, balanced [lo, hi) range of global chain ids owned by `rank`."""
    bounds = np.linspace(0, total_chains, world_size + 1).astype(int)
    return int(bounds[rank]), int(bounds[rank + 1])


def global_chain_seeds(base_seed, lo, hi):
    """uint64 Philox keys for global chains lo..hi-1."""
    return (np.uint64(base_seed) + np.arange(lo, hi, dtype=np.uint64)).astype(np.uint64)


def reduce_job_metrics(seconds, counts, device=None):
    """(max over ranks of each entry of `seconds`, sum over ranks of each entry of `counts`)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(seconds), dtype=torch.float64, device=device)
    c = torch.tensor(list(counts), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    return t.tolist(), c.tolist()


def combine_ess(per_rank_ess, device=None):
    """ESS of independent chain sets adds: sum the per-scalar vector over ranks, then take the min."""
    import torch
    import torch.distributed as dist
    v = torch.tensor(np.asarray(per_rank_ess, dtype="f8"), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(v, op=dist.ReduceOp.SUM)
    return float(v.min().item()), v.cpu().numpy()
