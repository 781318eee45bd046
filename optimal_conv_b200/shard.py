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


def _cpulist(txt):
    cpus = []
    for part in txt.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.extend(range(int(a), int(b or a) + 1))
    return cpus


def pin_rank_to_gpu_node(local_rank, world_size):
    """Bind this process to its share of the cores of the NUMA node its GPU is attached to (sysfs), so that the threads
    feeding the GPU and the pinned host buffers they first-touch are local to it and the ranks of one box do not fight
    over the same cores.  Returns what was done (for the bench record); never raises."""
    info = {"numa_node": None, "cpus": None}
    try:
        import torch
        pr = torch.cuda.get_device_properties(local_rank)
        bus = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = -1
        path = "/sys/bus/pci/devices/%s/numa_node" % bus
        if os.path.exists(path):
            node = int(open(path).read().strip())
        info["numa_node"] = node
        allowed = sorted(os.sched_getaffinity(0))
        cpus = allowed
        if node >= 0 and os.path.exists("/sys/devices/system/node/node%d/cpulist" % node):
            local = [c for c in _cpulist(open("/sys/devices/system/node/node%d/cpulist" % node).read()) if c in set(allowed)]
            if local:
                cpus = local
        # ranks whose GPUs share a node split its cores between them
        share = max(1, len(cpus) // max(1, world_size))
        mine = cpus[(local_rank * share) % len(cpus):][:share] if world_size > 1 else cpus
        if mine:
            os.sched_setaffinity(0, mine)
            info["cpus"] = "%d-%d (%d)" % (mine[0], mine[-1], len(mine))
    except Exception as e:  # placement is an optimisation, not a requirement
        info["error"] = str(e)[:80]
    return info
