"""Batch split of independent ciphertexts across ranks: no data-path
collective; the only collectives are a barrier and the max of the per-rank device times.
Pure torch.distributed plumbing so that it can be tested with gloo on CPU."""
import os


def world():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def my_units(n_units, rank, world_size):
    """Indices of the ciphertexts rank `rank` owns: dealt round-robin like images to GPUs."""
    return list(range(rank, n_units, world_size))


def max_over_ranks(values, device=None):
    """element-wise max of a list of floats over all ranks (identity when not distributed)"""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return list(values)
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def sum_over_ranks(values, device=None):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return list(values)
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.tolist()


def throughput(units_per_rank_per_step, steps, ms_per_rank, device=None):
    """whole-job units/s: all ranks' units over the slowest rank's device time"""
    total_units = sum_over_ranks([units_per_rank_per_step * steps], device)[0]
    ms = max_over_ranks([ms_per_rank], device)[0]
    return total_units / (ms / 1e3), ms
