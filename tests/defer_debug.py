"""GPU check of the deferred-transform plan against tests/defer_model.py, level by level (development aid)."""
import ctypes as C
import os
import sys

os.environ.setdefault("HEC_DEFER", "2")   # deferred whatever the batch (a single ciphertext here)

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import common  # noqa: E402
import defer_model as dm  # noqa: E402
from optimal_conv_b200 import hec, params as PR, synth  # noqa: E402
from oracle.orc import Oracle  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
norm = int(sys.argv[2]) if len(sys.argv) > 2 else 1
N = 1 << PR.LOGN
o = Oracle(PR.LOGN, common.Q2, common.P1)
idx_np = o.monomial_pts()
w = synth.conv_workload(common.Q2, common.P1, PR.LOGN, B, seed=900 + B)
trace = []
want = dm.model(o, w, norm, PR.SCALE, idx_np, trace=trace)
c = hec.Context(PR.LOGN, common.Q2, common.P1)
G = common.GpuConv(c, w, idx_np, norm=norm)
plan = c.plan(G.ker, norm, PR.SCALE, PR.SCALE, G.idx, G.bias, 1)
print("deferred:", plan.deferred)
outs = plan.run(G.cts)
g0, g1 = outs[0].download()
print("final c0", np.array_equal(g0[0], want[0]), "c1", np.array_equal(g1[0], want[1]))
L = c.L
L.hec_plan_debug_level.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
na = B // norm
for lv, cts in enumerate(trace):
    n = na >> lv
    for half, nm in ((0, "U"), (1, "e")):
        if lv == 0 and half == 0 and len(trace) > 1:
            continue    # the U halves of stage A are not materialised when a pack level follows
        buf = np.empty((n, 2, N), dtype=np.uint64)
        rc = L.hec_plan_debug_level(plan.h, lv, half, buf.ctypes.data_as(C.c_void_p))
        assert rc == 0, rc
        for u in range(n):
            for p in range(2):
                ref = np.asarray(cts[u][p][half], dtype=np.uint64)
                ok = np.array_equal(buf[u, p], ref)
                if not ok:
                    bad = np.nonzero(buf[u, p] != ref)[0]
                    print("level", lv, nm, "ct", u, "poly", p, "MISMATCH", len(bad), "first", bad[:8], "got", buf[u, p][bad[:3]], "want", ref[bad[:3]])
                else:
                    print("level", lv, nm, "ct", u, "poly", p, "ok")
