"""Generates tests/golden/conv_golden.json from the CPU oracle (run from the repo root:
`python tests/golden/make_golden.py`): SHA-256 of oracle outputs on seeded inputs at N = 2^16.
The oracle itself is pinned against the reference's compiled code (make_ref_vectors.py,
make_ref_eval_vectors.py); the B4_norm1 entry here is additionally checked to equal the digest the
reference's own main.conv_then_pack produced at N = 2^16 (tests/test_ref_eval_vectors.py)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import common  # noqa: E402
from optimal_conv_b200 import params as PR, synth  # noqa: E402
from oracle.orc import Ct, Oracle  # noqa: E402

out = {}
o = Oracle(PR.LOGN, common.Q2, common.P1)
idx = o.monomial_pts()
out["monomials"] = common.sha(idx)
conv = {}
for cfg in common.GOLDEN_CONFIGS:
    w = common.workload(cfg)
    r = common.oracle_conv(o, w, cfg["norm"], float(1 << cfg["out_log"]), idx)
    conv[cfg["name"]] = {"c0": common.sha(r.c0), "c1": common.sha(r.c1), "scale": r.scale}
out["conv"] = conv
# ring level: NTT / InvNTT of a seeded limb for every modulus of sets 6 and 7 and the 5 P
allq = PR.Q_SET6 + [PR.Q_SET7[1], PR.Q_SET7[13]]
oa = Oracle(PR.LOGN, allq, PR.P_ALL)
ntt = {}
for ring, mods in ((0, allq), (1, PR.P_ALL)):
    for limb, q in enumerate(mods):
        a = synth.uniform_mod(500 + limb + 100 * ring, 1 << PR.LOGN, q)
        ntt["%d:%x" % (ring, q)] = {"fwd": common.sha(oa.ntt(a, limb, ring)), "inv": common.sha(oa.intt(a, limb, ring))}
out["ntt"] = ntt
# rotation (RotateGal) at level 0 and 1 with the pack key for galEl 2^13+1
w = common.workload(common.GOLDEN_CONFIGS[1])
g = (1 << 13) + 1
rot = {}
for lv in (0, 1):
    ct = Ct(w["ct"][0][0][:lv + 1], w["ct"][0][1][:lv + 1], PR.SCALE)
    r = o.rotate_gal(ct, g, w["keys"][12])
    rot["level%d" % lv] = {"c0": common.sha(r.c0), "c1": common.sha(r.c1)}
out["rotate_gal_2^13+1"] = rot
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "conv_golden.json")
json.dump(out, open(path, "w"), indent=1, sort_keys=True)
print("wrote", path)
