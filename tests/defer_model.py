"""CPU model of the deferred-transform dataflow of the fused conv path, checked against the oracle
(test infrastructure: tests/test_defer_model.py runs it; tests/defer_debug.py compares a GPU plan with it level by level).

Every level-0 polynomial of the pack tree is carried as a pair (U, e): value = U - NTT(e) mod q0, U in the NTT domain,
e in the coefficient domain.  The forward transforms the rescale (stage A) and the mod-down (stage B) end with are not
run; monomial products, sums and the automorphism act on e in the coefficient domain (shift / add / index map with sign),
and NTT(e) is formed only where the VALUE is needed: for the c1 polynomial that enters a key switch (as
d = InvNTT(U) - e, which is also the digit) and once per output polynomial at the end.  All of it is exact arithmetic
modulo q0, so the canonical result is the reference's, bit for bit; what goes through the special prime is unchanged.

python tests/defer_model.py [B] [norm]   -> compares with Oracle.conv_then_pack at N = 2^16
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from optimal_conv_b200 import params as PR, synth  # noqa: E402
from oracle.orc import Oracle  # noqa: E402
import common  # noqa: E402

N = 1 << PR.LOGN
R = 1 << 64


def O(a):
    return np.asarray(a, dtype=object)


def U64(a):
    return np.asarray(a, dtype=np.uint64)


def shift_neg(e, s, q):
    """coefficients of X^s * e(X) in Z_q[X]/(X^N+1)"""
    out = np.empty_like(e)
    out[s:] = e[:N - s]
    out[:s] = (q - e[N - s:]) % q
    return out


def sigma_coef(e, g, q):
    """coefficients of e(X^g): X^n -> X^(n g mod 2N), sign flips past N (ring.Permute)"""
    n = np.arange(N, dtype=np.int64)
    m = (n * g) % (2 * N)
    out = np.empty_like(e)
    pos = m < N
    out[m[pos]] = e[pos]
    out[m[~pos] - N] = (q - e[~pos]) % q
    return out


def float_v(y, p):
    return np.array([int(float(int(v)) / float(p)) for v in y], dtype=object)


def model(o, w, norm, out_scale, idx_np, bias=True, trace=None):
    Q, P = o.Q, o.P
    q0, q1, p0 = Q[0], Q[1], P[0]
    B = w["B"]
    na = B // norm
    # SetScale bookkeeping as hec_plan_create does it
    target = out_scale / (B // norm)
    s = PR.SCALE * PR.SCALE
    k, up = o.const_limbs(1, target / s)
    k = [int(x) for x in k]
    c0, c1 = w["ct"][0]
    half1 = (q1 - 1) >> 1
    q1inv = pow(q1 % q0, -1, q0)
    # ---- stage A: (U, e) per active channel and polynomial
    cts = []
    for a in range(na):
        pt = w["pt_ker"][a * norm]
        pair = []
        for poly in (c0, c1):
            x0 = O(poly[0]) * O(pt[0]) % q0 * k[0] % q0
            x1 = O(poly[1]) * O(pt[1]) % q1 * k[1] % q1
            t = O(o.intt(U64(x1), 1))
            t = (t + half1) % q1
            r = (t % q0 - half1 % q0) % q0                     # centred remainder, lifted
            pair.append((x0 * q1inv % q0, r * q1inv % q0))     # value = (x0 - NTT(r)) / q1
        cts.append(pair)
    if trace is not None:
        trace.append(cts)
    # ---- stage B
    pinv = pow(p0 % q0, -1, q0)
    step = B // 2
    log_step = step.bit_length() - 1
    j = PR.LOGN - log_step
    n = na
    while n > 1:
        g = (1 << j) + 1
        mono = O(idx_np[log_step])
        key = w["keys"][j - 1]      # [digit][poly][limb q0,q1,p0][N], Montgomery form
        kq = [O(key[0][c][0]) * pow(R, -1, q0) % q0 for c in range(2)]
        kp = [O(key[0][c][2]) * pow(R, -1, p0) % p0 for c in range(2)]
        perm = o.permute_index(g)
        nxt = []
        for u in range(n // 2):
            a, b = cts[u], cts[u + n // 2]
            # tmp2 = a - b X^step, tmp1 = a + b X^step, on both halves of the representation
            mU = [b[c][0] * mono % q0 for c in range(2)]
            mE = [shift_neg(b[c][1], step, q0) for c in range(2)]
            U2 = [(a[c][0] - mU[c]) % q0 for c in range(2)]
            E2 = [(a[c][1] - mE[c]) % q0 for c in range(2)]
            U1 = [(a[c][0] + mU[c]) % q0 for c in range(2)]
            E1 = [(a[c][1] + mE[c]) % q0 for c in range(2)]
            # the digit: coefficients of tmp2.c1
            d = (O(o.intt(U64(U2[1]), 0)) - E2[1]) % q0
            dQ = O(o.ntt(U64(d), 0))                           # its NTT form under q0 (the value of tmp2.c1)
            dP = O(o.ntt(U64(d % p0), 0, ring=1))
            out = []
            for c in range(2):
                accQ = dQ * kq[c] % q0
                accP = dP * kp[c] % p0
                y = O(o.intt(U64(accP), 0, ring=1))
                v = float_v(y, p0)
                ext = (y % q0 - v * (p0 % q0)) % q0            # exact basis extension P -> q0
                Uk = accQ * pinv % q0                          # key-switch output = (accQ - NTT(ext)) / P
                Ek = ext * pinv % q0
                if c == 0:
                    Uk = (Uk + U2[0]) % q0
                    Ek = (Ek + E2[0]) % q0
                Un = (U1[c] + Uk[perm]) % q0                   # sigma_g in the NTT domain: out[i] = in[index_g[i]]
                En = (E1[c] + sigma_coef(Ek, g, q0)) % q0
                out.append((Un, En))
            nxt.append(out)
        cts = nxt
        if trace is not None:
            trace.append(cts)
        n //= 2
        step //= 2
        log_step -= 1
        j += 1
    # ---- the one transform per output polynomial
    res = []
    for c in range(2):
        Ufin, Efin = cts[0][c]
        v = (Ufin - O(o.ntt(U64(Efin), 0))) % q0
        if c == 0 and bias:
            v = (v + O(w["bias"])) % q0
        res.append(U64(v))
    return res


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    norm = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    o = Oracle(PR.LOGN, common.Q2, common.P1)
    idx_np = o.monomial_pts()
    w = synth.conv_workload(common.Q2, common.P1, PR.LOGN, B, seed=900 + B)
    ref = common.oracle_conv(o, w, norm, PR.SCALE, idx_np)
    got = model(o, w, norm, PR.SCALE, idx_np)
    ok0, ok1 = np.array_equal(got[0], ref.c0[0]), np.array_equal(got[1], ref.c1[0])
    print("B", B, "norm", norm, "c0", ok0, "c1", ok1)
    return 0 if ok0 and ok1 else 1


if __name__ == "__main__":
    sys.exit(main())
