"""Golden vectors produced by the reference's OWN compiled code.

Run from the repo root in the build container (needs /root/reference/test_run, objdump, nm):
    python tests/golden/make_ref_vectors.py
The prebuilt reference binary is never executed: its pure ring routines are interpreted from their
disassembly by tests/golden/x86emu.py on a private register file and memory (no syscalls, no I/O).
Routines covered (symbols of github.com/dwkim606/test_lattigo/ring in test_run):
  primitiveRoot (+ ModExp, BRed)  -> the generator g every NTT table is built from, all 35 moduli
  NTTLazy / NTT / InvNTT / InvNTTLazy -> transforms of seeded inputs with the tables built from g
  PermuteNTTIndex                 -> automorphism index tables
  reconstructRNS + multSum        -> exact basis extension incl. the float64 overflow count v
Hooks (python replacements of callees that need the Go runtime / math/big): runtime.makeslice
(allocation), ring.getFactors (prime factors of q-1: pure number theory), ring.BRedParams
(floor(2^128/q)).  Output: tests/golden/ref_vectors.json, consumed by tests/test_ref_vectors.py.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from x86emu import Emu, M64  # noqa: E402
from optimal_conv_b200 import params as PR, synth  # noqa: E402

BIN = "/root/reference/test_run"
P = "github.com/dwkim606/test_lattigo/ring."
R = 1 << 64


def sha(vals):
    return hashlib.sha256(np.array(vals, dtype=np.uint64).tobytes()).hexdigest()


def factors(n):
    out, p = [], 2
    while p * p <= n:
        if n % p == 0:
            out.append(p)
            while n % p == 0:
                n //= p
        p += 1 if p == 2 else 2
    if n > 1:
        out.append(n)
    return out


def new_emu():
    e = Emu(BIN)
    for f in ["NTTLazy", "NTT", "InvNTT", "InvNTTLazy", "PermuteNTTIndex", "primitiveRoot", "ModExp", "BRed",
              "reconstructRNS", "multSum"]:
        e.load(P + f)

    def h_getfactors(em):
        sp = em.r[4]
        fs = factors(em.rq(sp))
        a = em.alloc(8 * len(fs))
        em.write_u64s(a, fs)
        em.wq(sp + 8, a); em.wq(sp + 16, len(fs)); em.wq(sp + 24, len(fs))

    def h_bredparams(em):
        sp = em.r[4]
        q = em.rq(sp)
        b = (1 << 128) // q
        a = em.alloc(16)
        em.write_u64s(a, [b >> 64, b & M64])
        em.wq(sp + 8, a); em.wq(sp + 16, 2); em.wq(sp + 24, 2)

    def h_makeslice(em):
        sp = em.r[4]
        esize = em.load_const(em.rq(sp))  # runtime._type.size is the first word
        a = em.alloc(esize * em.rq(sp + 16))
        em.wq(sp + 24, a)

    e.hook(P + "getFactors", h_getfactors)
    e.hook(P + "BRedParams", h_bredparams)
    e.hook("runtime.makeslice", h_makeslice)
    return e


def tables(q, g, logN):
    """NttPsi / NttPsiInv (Montgomery form, index brev(j)) from the generator g, as the oracle builds them"""
    N = 1 << logN
    psi = pow(g, (q - 1) // (2 * N), q)
    psii = pow(psi, -1, q)
    t, ti = [0] * N, [0] * N
    a = b = 1
    for j in range(N):
        r = int(format(j, "0%db" % logN)[::-1], 2)
        t[r], ti[r] = a * R % q, b * R % q
        a, b = a * psi % q, b * psii % q
    return t, ti


# ---- reconstructRNS + multSum: exact basis extension with the float64 overflow count ----
# sources S (digit limbs), one target modulus; tables as Lattigo's genModUpParams builds them
def modup_case(e, S, target, cols):
    n = len(S)
    PP, V, Y, QA, QI, QB, RES, QPJ, QISP = (e.alloc(4096) for _ in range(9))
    rows = [e.alloc(8 * 8) for _ in range(n)]
    for i in range(n):
        e.write_u64s(rows[i], [c[i] for c in cols])
        e.write_u64s(PP + 24 * i, [rows[i], 8, 8])
    Ys = [e.alloc(32 * 8) for _ in range(8)]
    Qd = 1
    for s in S:
        Qd *= s
    e.write_u64s(QA, S)
    e.write_u64s(QI, [pow(s, -1, R) for s in S])
    e.write_u64s(QB, [pow(Qd // s, -1, s) * R % s for s in S])
    e.call(P + "reconstructRNS", [n, 0, PP, n, n, V] + Ys + [QA, n, n, QI, n, n, QB, n, n])
    v = e.read_u64s(V, 8)
    e.write_u64s(QPJ, [(-k * Qd) % target for k in range(n + 1)])
    e.write_u64s(QISP, [(Qd // s) % target * R % target for s in S])
    e.call(P + "multSum", [RES, V] + Ys + [n, target, pow(target, -1, R), QPJ, n + 1, n + 1, QISP, n, n])
    res = e.read_u64s(RES, 8)
    return v, [r % target for r in res], res


def main():
    out = {"binary": "test_run (go1.16.6, test_lattigo v0.0.0-20220812213541-eb33b0555aaa)", "generator": {}, "ntt": {},
           "permute_index": {}, "modup": []}
    e = new_emu()
    allq = sorted(set(PR.Q_SET6 + PR.Q_SET7 + PR.P_ALL))
    # ---- primitiveRoot for every modulus on the path ----
    for q in allq:
        g = e.call(P + "primitiveRoot", [q, 0])[1]
        out["generator"]["%x" % q] = g
    # ---- transforms ----
    IN, OUT, PSI, BRED = 0x10000000, 0x20000000, 0x30000000, 0x40000000
    cases = [(PR.Q_SET6[0], 8), (PR.Q_SET6[1], 8), (PR.P_ALL[0], 8), (PR.Q_SET6[5], 8), (PR.Q_SET7[1], 8),
             (PR.Q_SET6[0], 16), (PR.P_ALL[0], 16)]
    for q, logN in cases:
        N = 1 << logN
        g = out["generator"]["%x" % q]
        t, ti = tables(q, g, logN)
        qinv = pow(q, -1, R)
        b = (1 << 128) // q
        ninv = pow(N, -1, q) * R % q
        a = synth.uniform_mod(500 + logN, N, q)
        e.write_u64s(IN, a)
        e.write_u64s(BRED, [b >> 64, b & M64])
        rec = {}
        e.write_u64s(PSI, t)
        e.call(P + "NTTLazy", [IN, N, N, OUT, N, N, N, PSI, N, N, q, qinv, BRED, 2, 2])
        lazy = e.read_u64s(OUT, N)
        rec["ntt_lazy_max_over_q"] = max(lazy) / q
        rec["ntt_lazy_canonical"] = sha([v % q for v in lazy])
        e.call(P + "NTT", [IN, N, N, OUT, N, N, N, PSI, N, N, q, qinv, BRED, 2, 2])
        rec["ntt"] = sha(e.read_u64s(OUT, N))
        e.write_u64s(PSI, ti)
        e.call(P + "InvNTT", [IN, N, N, OUT, N, N, N, PSI, N, N, ninv, q, qinv])
        rec["intt"] = sha(e.read_u64s(OUT, N))
        e.call(P + "InvNTTLazy", [IN, N, N, OUT, N, N, N, PSI, N, N, ninv, q, qinv])
        lz = e.read_u64s(OUT, N)
        rec["intt_lazy_max_over_q"] = max(lz) / q
        rec["intt_lazy_canonical"] = sha([v % q for v in lz])
        out["ntt"]["%x:%d" % (q, logN)] = rec
        print("ntt", hex(q), logN, rec["ntt_lazy_max_over_q"], e.steps, flush=True)
    # ---- PermuteNTTIndex ----
    for logN, gal in [(16, (1 << 13) + 1), (16, (1 << 16) + 1), (16, 5), (16, pow(5, 2 * 65536 - 3, 2 * 65536)), (10, 25)]:
        N = 1 << logN
        res = e.call(P + "PermuteNTTIndex", [gal, N, 0, 0, 0])
        idx = e.read_u64s(res[2], N)
        out["permute_index"]["%d:%d" % (logN, gal)] = sha(idx)
    p0, q0 = PR.P_ALL[0], PR.Q_SET6[0]
    edge = [[p0 - 1], [p0 - 129], [p0 - 130], [p0 - 64], [0], [1], [p0 // 2], [12345678901234567]]
    v, res, raw = modup_case(e, [p0], q0, edge)
    out["modup"].append({"S": ["%x" % p0], "target": "%x" % q0, "cols": [[str(x) for x in c] for c in edge],
                         "v": v, "res_mod_target": [str(x) for x in res], "raw_max_over_target": max(raw) / q0})
    S2 = PR.Q_SET7[:2]
    cols2 = [[int(synth.uniform_mod(70 + k, 1, S2[0])[0]), int(synth.uniform_mod(80 + k, 1, S2[1])[0])] for k in range(8)]
    cols2[0] = [S2[0] - 1, S2[1] - 1]
    cols2[1] = [0, 0]
    for tgt in PR.P_PACK_BL:
        v, res, raw = modup_case(e, S2, tgt, cols2)
        out["modup"].append({"S": ["%x" % s for s in S2], "target": "%x" % tgt, "cols": [[str(x) for x in c] for c in cols2],
                             "v": v, "res_mod_target": [str(x) for x in res], "raw_max_over_target": max(raw) / tgt})
    S5 = PR.Q_SET6[:5]
    cols5 = [[int(synth.uniform_mod(90 + 10 * k + i, 1, S5[i])[0]) for i in range(5)] for k in range(8)]
    cols5[0] = [s - 1 for s in S5]
    v, res, raw = modup_case(e, S5, PR.P_ALL[2], cols5)
    out["modup"].append({"S": ["%x" % s for s in S5], "target": "%x" % PR.P_ALL[2], "cols": [[str(x) for x in c] for c in cols5],
                         "v": v, "res_mod_target": [str(x) for x in res], "raw_max_over_target": max(raw) / PR.P_ALL[2]})
    out["interpreted_instructions"] = e.steps
    json.dump(out, open(os.path.join(HERE, "ref_vectors.json"), "w"), indent=1, sort_keys=True)
    print("wrote ref_vectors.json;", e.steps, "instructions interpreted")


if __name__ == "__main__":
    main()
