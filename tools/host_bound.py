"""Is the generic path host-bound?  Wall time until hec_bootstrap_ctos RETURNS (launches enqueued) against the time until
the stream is idle, for a single ciphertext (development aid)."""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from optimal_conv_b200 import hec, params as PR, synth  # noqa: E402

N = 1 << PR.LOGN
Q, P = PR.Q_SET6, PR.P_ALL
ctx = hec.Context(PR.LOGN, Q, P)
keys, kconj, rlk, b = synth.ctos_operands(N)
for r, k in keys.items():
    ctx.upload_swk(ctx.galois_for_rotation(r), k, 27)
ctx.upload_swk(2 * N - 1, kconj, 27)
ctx.upload_rlk(rlk, 27)
mats = [ctx.upload_ptdiag(PR.LOGN - 1, n1, ml, ms, D) for D, n1, ml, ms in b["mats"]]
a0, a1 = synth.uniform_limbs(61, Q[:2], N), synth.uniform_limbs(62, Q[:2], N)
A = ctx.upload_ct(a0, a1, PR.SCALE * 2.0 ** 8)
ret, tot = [], []
for i in range(8):
    ctx.sync()
    l0 = ctx.launch_count()
    t0 = time.perf_counter()
    g0, g1, _ = ctx.BootstrappConv_CtoS(A, b, mats)
    t1 = time.perf_counter()
    ctx.sync()
    t2 = time.perf_counter()
    n = ctx.launch_count() - l0
    g0.free()
    if g1 is not None:
        g1.free()
    ret.append(1e3 * (t1 - t0)); tot.append(1e3 * (t2 - t0))
ret.sort(); tot.sort()
print("launches", n, "call returns after %.2f ms, stream idle after %.2f ms (medians)" % (ret[len(ret) // 2], tot[len(tot) // 2]))
