import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
from optimal_conv_b200 import hec, params as PR, synth
N = 1 << 16
Q, P = PR.Q_SET6[:16], PR.P_ALL
c = hec.Context(16, Q, P)
rlk = np.stack([np.stack([synth.uniform_limbs(8000 + 10 * d + k, Q + P, N) for k in range(2)]) for d in range(4)])
c.upload_rlk(rlk, 15)
A = c.upload_ct(synth.uniform_limbs(61, Q, N), synth.uniform_limbs(62, Q, N), PR.SCALE)
for _ in range(3): c.evalReLU(A, 0.0, PR.SCALE).free()
c.sync()
for _ in range(3):
    t0 = time.perf_counter(); r = c.evalReLU(A, 0.0, PR.SCALE); t1 = time.perf_counter(); c.sync(); t2 = time.perf_counter(); r.free()
    print("enqueue %.2f ms, total %.2f ms" % ((t1 - t0) * 1e3, (t2 - t0) * 1e3))
