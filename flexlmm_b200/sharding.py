"""SNP-block sharding across GPUs (one process per GPU) — host-side logic only.

The reference scatters FIT_MODEL over chromosome x phenotype x permutation seed as independent R processes
(subworkflows/local/lmm.nf:42-71) and reduces with `cat` + aggregate(min) (get_null_p_dist.nf:24-25,77).  Here SNPs are
the independent unit: each rank gets a contiguous block of var_idx, per-SNP outputs are disjoint, and only the length-P
min-p vector (min) and the tested-variant count (sum) cross ranks.
"""
import numpy as np


def shard_bounds(M, rank, world):
    """Contiguous, balanced split of M SNPs: ranks < M % world get one extra."""
    base, extra = divmod(int(M), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_var_idx(var_idx, rank, world):
    lo, hi = shard_bounds(len(var_idx), rank, world)
    return np.asarray(var_idx)[lo:hi]


def allreduce_minp(min_pval, nvars, dist=None):
    """min over ranks of min_pval, sum of nvars, via torch.distributed (gloo on CPU in tests; the CUDA path does the same
    reduction with ncclAllReduce inside libflexlmm_cuda.so when a communicator is attached)."""
    import torch
    mp = torch.as_tensor(np.asarray(min_pval, dtype=np.float64)).clone()
    nv = torch.tensor([int(nvars)], dtype=torch.int64)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(mp, op=dist.ReduceOp.MIN)
        dist.all_reduce(nv, op=dist.ReduceOp.SUM)
    return mp.numpy(), int(nv[0])


def gather_per_snp(local, M, rank, world, dist=None):
    """Concatenate per-SNP arrays of all ranks in var_idx order (keeps the TSV row order of fit_model.nf:145-169)."""
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return np.asarray(local)
    parts = [None] * world
    dist.all_gather_object(parts, np.asarray(local))
    return np.concatenate(parts, axis=0)
