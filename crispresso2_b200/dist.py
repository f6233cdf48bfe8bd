"""Multi-GPU plumbing: reads shard embarrassingly across ranks (one process per GPU, torch.distributed);
the only exchange on the path is one all-reduce (sum, int64) of the per-position count block at the end --
the merge CRISPRessoCORE.py performs implicitly by looping over one variantCache (:3964-4115).
"""
import numpy as np


def shard_bounds(n_items, rank, world):
    """Contiguous equal slices, like the reference's own sharding of unique reads
    (get_variant_cache_equal_boundaries, CRISPRessoCORE.py:1172-1195)."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_by_cells(lengths, rank, world):
    """Mixed-length batches: deal reads so every rank gets about the same number of DP cells (I*J, I fixed)."""
    order = np.argsort(-np.asarray(lengths), kind="stable")
    return np.sort(order[rank::world])


class _CudaView:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 2}


def allreduce_counts(engine, group=None):
    """Sums the engines' count blocks over all ranks.  On CUDA the NCCL all-reduce runs in place on the engine's
    device block (zero copy, NVLink); on the CPU test path (gloo) the block is reduced through a host tensor.
    -> reduced int64 numpy array (every rank)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return engine.counts_raw()
    backend = dist.get_backend(group)
    if backend == "nccl":
        ptr, n = engine.counts_device()
        engine.sync()
        t = torch.as_tensor(_CudaView(ptr, n), device="cuda:%d" % engine.device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        torch.cuda.synchronize(engine.device)
        return engine.counts_raw()
    t = torch.from_numpy(engine.counts_raw())
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.numpy()


def allreduce_array(arr, engine, group=None):
    """Sum of an int64 host array over the ranks (NCCL through a tensor on the engine's GPU, gloo on the host)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return arr
    if dist.get_backend(group) == "nccl":
        t = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.int64)).to("cuda:%d" % engine.device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        return t.cpu().numpy()
    t = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.int64).copy())
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.numpy()
