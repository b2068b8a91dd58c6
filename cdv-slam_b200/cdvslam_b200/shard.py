"""Multi-GPU plumbing of the BA hot path: independent windows are partitioned over ranks (replicas only -- one
sliding-window BA does not shard); there is NO data-path collective.  torch.distributed is used for
the barrier around the timed region and for the max-over-ranks of the device time only (backend nccl on GPUs, gloo in
the CPU tests)."""
import torch


def windows_of_rank(n_windows, world, rank):
    """Round-robin assignment of independent windows / sequences to ranks (BASELINE config c5)."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    return list(range(rank, n_windows, world))


def max_over_ranks(value, device="cpu", dist=None):
    """max of a python float over all ranks (identity without a process group)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device="cpu", dist=None):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def aggregate_throughput(units_this_rank, ms_this_rank, device="cpu", dist=None):
    """Whole-job throughput = units processed by all ranks / max-over-ranks time (the bench contract)."""
    total = sum_over_ranks(units_this_rank, device, dist)
    ms = max_over_ranks(ms_this_rank, device, dist)
    return total / (ms * 1e-3), ms
