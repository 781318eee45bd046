"""device time of one plan run (B = 16) for small batches, deferred transforms on / off (HEC_DEFER): where the crossover is"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from optimal_conv_b200 import hec, params as PR, synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
Q2, P1 = PR.Q_SET6[:2], PR.P_ALL[:1]
c = hec.Context(PR.LOGN, Q2, P1)
N = 1 << PR.LOGN
w = synth.conv_workload(Q2, P1, PR.LOGN, B, seed=1, n_ct=1)
# monomials: NTT of X^(2^i) through the library itself
idx = []
for i in range(PR.LOGN):
    m = np.zeros(N, dtype=np.uint64)
    m[1 << i] = 1
    idx.append(c.upload_pt(c.ntt(m, 0)[None, :], 1.0))
ker = [c.upload_pt(w["pt_ker"][i], PR.SCALE) for i in range(B)]
bias = c.upload_pt(w["bias"][None, :], PR.SCALE)
for j, k in w["keys"].items():
    c.upload_swk((1 << (j + 1)) + 1, k, 0)
for M in (1, 2, 4, 8, 16, 32):
    cts = [c.upload_ct(w["ct"][0][0], w["ct"][0][1], PR.SCALE) for _ in range(M)]
    plan = c.plan(ker, 1, PR.SCALE, PR.SCALE, idx, bias, M)
    for _ in range(5):
        plan.run(cts)
    torch.cuda.synchronize()
    ts = []
    for _ in range(30):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        plan.run(cts)
        c.sync()
        ts.append(1e3 * (time.perf_counter() - t0))
    ts.sort()
    print("B", B, "M", M, "deferred", plan.deferred, "ms/run (median wall, synced)", round(ts[len(ts) // 2], 4))
    plan.destroy()
    for x in cts:
        x.free()
c.close()
