"""CPU, world_size 2 over gloo: the N>1 plumbing of bench.py (batch split of independent
ciphertexts, max-over-ranks timing, whole-job throughput) with the oracle standing in for
the device so that every rank really computes its own shard."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import torch.distributed as dist
    import common
    from optimal_conv_b200 import params as PR, shard
    from oracle.orc import Oracle
    dist.init_process_group("gloo", rank=rank, world_size=world)
    r, _, w = shard.world()
    assert (r, w) == (rank, world)
    n_units = 5  # ragged: rank 0 gets 3 ciphertexts, rank 1 gets 2
    mine = shard.my_units(n_units, rank, world)
    cfg = {"B": 4, "seed": 301}
    wl = common.workload(cfg, n_ct=n_units)
    o = Oracle(PR.LOGN, common.Q2, common.P1)
    idx = o.monomial_pts()
    hashes = {m: common.sha(common.oracle_conv(o, wl, 1, PR.SCALE, idx, m=m).c0) for m in mine}
    ms = 10.0 * (rank + 1)  # pretend device times: the slowest rank defines the step
    val, ms_max = shard.throughput(len(mine), steps=2, ms_per_rank=ms)
    gathered = [None] * world
    dist.all_gather_object(gathered, hashes)
    dist.barrier()
    if rank == 0:
        q.put((val, ms_max, gathered))
    dist.destroy_process_group()


def test_two_ranks_split_independent_ciphertexts():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import common
    from optimal_conv_b200 import params as PR, shard
    from oracle.orc import Oracle
    assert shard.my_units(5, 0, 2) == [0, 2, 4] and shard.my_units(5, 1, 2) == [1, 3]
    assert shard.my_units(0, 0, 2) == [] and shard.my_units(1, 1, 2) == []
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    val, ms_max, gathered = q.get(timeout=180)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    # whole-job throughput = all units of all ranks / slowest rank's time
    assert ms_max == 20.0 and abs(val - (3 + 2) * 2 / 0.020) < 1e-6
    # every ciphertext was processed exactly once and sharding does not change any bit
    merged = {}
    for g in gathered:
        merged.update(g)
    assert sorted(merged) == [0, 1, 2, 3, 4]
    o = Oracle(PR.LOGN, common.Q2, common.P1)
    wl = common.workload({"B": 4, "seed": 301}, n_ct=5)
    idx = o.monomial_pts()
    for m in (0, 3):
        assert merged[m] == common.sha(common.oracle_conv(o, wl, 1, PR.SCALE, idx, m=m).c0)
