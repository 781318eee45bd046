"""Soak: repeat the fused conv plan (and a generic rotation) many times on fresh and reused buffers and require
identical digests every time -- a cheap detector for rare races / uninitialised reads.  python tools/soak.py [iters]"""
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import common  # noqa: E402
from optimal_conv_b200 import hec, params as PR, synth  # noqa: E402
from oracle.orc import Oracle  # noqa: E402  (only for the monomial plaintexts and one reference result)

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 300
c = hec.Context(PR.LOGN, common.Q2, common.P1)
o = Oracle(PR.LOGN, common.Q2, common.P1)
idx = o.monomial_pts()
w = common.workload({"B": 16, "seed": 5}, n_ct=8)
G = common.GpuConv(c, w, idx)
plan = c.plan(G.ker, 1, PR.SCALE, PR.SCALE, G.idx, G.bias, 8)
ref = [common.oracle_conv(o, w, 1, PR.SCALE, idx, m=m) for m in range(8)]
want = hashlib.sha256(b"".join(np.ascontiguousarray(r.c0).tobytes() + np.ascontiguousarray(r.c1).tobytes() for r in ref)).hexdigest()
bad = 0
for it in range(iters):
    outs = plan.run(G.cts)
    h = hashlib.sha256()
    for r in outs:
        g0, g1 = r.download()
        h.update(g0.tobytes() + g1.tobytes())
    bad += h.hexdigest() != want
print("fused plan: %d iterations, %d mismatches" % (iters, bad))
g = (1 << 13) + 1
ct = plan.run(G.cts)[0]   # a level-0 ciphertext: the pack keys are uploaded for level 0 only
first = None
for it in range(iters):
    out = c.CopyNew(ct)
    c.RotateGal(ct, g, out)
    g0, g1 = out.download()
    d = hashlib.sha256(g0.tobytes() + g1.tobytes()).hexdigest()
    first = first or d
    bad += d != first
    out.free()
print("generic RotateGal: %d iterations, total mismatches %d" % (iters, bad))
sys.exit(1 if bad else 0)
