"""Soak of the generic path (programmatic dependent launch on): repeat evalReLU, a rotation and a relinearised
multiplication many times, with ciphertexts allocated and freed in between so that pool memory is reused, and require
identical digests every time.  python tools/soak_generic.py [iters]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import common  # noqa: E402
from optimal_conv_b200 import hec, params as PR, synth  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 100
N, level = 1 << PR.LOGN, 15
Q, P = PR.Q_SET6[:level + 1], PR.P_ALL
c = hec.Context(PR.LOGN, Q, P)
beta = (level + 1 + len(P) - 1) // len(P)
key = lambda s: np.stack([np.stack([synth.uniform_limbs(s + 10 * d + k, Q + P, N) for k in range(2)]) for d in range(beta)])
c.upload_rlk(key(8000), level)
g = c.galois_for_rotation(5)
c.upload_swk(g, key(9000), level)
A = c.upload_ct(synth.uniform_limbs(61, Q, N), synth.uniform_limbs(62, Q, N), PR.SCALE)
B = c.upload_ct(synth.uniform_limbs(63, Q, N), synth.uniform_limbs(64, Q, N), PR.SCALE)


def digest(ct):
    a, b = ct.download()
    return common.sha(a, b)


want, bad = None, 0
for it in range(iters):
    r = c.evalReLU(A, 0.0, PR.SCALE)
    rot = c.CopyNew(B)
    c.RotateGal(B, g, rot)
    m = c.MulRelinNew(A, rot)
    c.Rescale(m, PR.SCALE)
    got = (digest(r), digest(rot), digest(m))
    for x in (r, rot, m):
        x.free()
    if want is None:
        want = got
    elif got != want:
        bad += 1
        print("iteration", it, "differs")
print("soak_generic: %d iterations, %d mismatches" % (iters, bad))
c.close()
sys.exit(1 if bad else 0)
