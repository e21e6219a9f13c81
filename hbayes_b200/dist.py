"""Multi-GPU plumbing for the evidence-row sharding: rows are independent, every rank holds the
same compiled program and takes a contiguous slice of the rows; there is NO data-path collective.  The only
communication is the timing protocol of bench.py (barrier + max over ranks), done with torch.distributed
(NCCL on the GPUs, gloo in the CPU tests)."""
from __future__ import annotations


def shard_rows(n_rows: int, rank: int, world: int):
    """Contiguous [lo, hi) slice of rank `rank`; slices differ by at most one row and tile [0, n_rows)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(n_rows, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def max_over_ranks(value: float, device=None) -> float:
    """max of a per-rank scalar (device time in ms); identity when torch.distributed is not initialised."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def whole_job_rate(rows_per_rank, steps: int, ms_max: float) -> float:
    """units all ranks processed / max-over-ranks time (bench.py's `value`)."""
    return float(sum(rows_per_rank)) * steps / (ms_max / 1e3)
