"""Golden vectors from the reference's OWN compiled evaluator code, up to main.conv_then_pack itself.

Run from the repo root in the build container (needs /root/reference/test_run, objdump, nm):
    python tests/golden/make_ref_eval_vectors.py            # small-N cases, ~2 min
    python tests/golden/make_ref_eval_vectors.py --only n8_B4      # regenerate one small conv case
    python tests/golden/make_ref_eval_vectors.py --evalops | --relu | --lt | --ctos | --btp | --ring | --conv | --encode | --hostprep  # regenerate one group
    python tests/golden/make_ref_eval_vectors.py --full B4_norm1   # N = 2^16 golden config, ~1 h
The prebuilt binary is never executed.  tests/golden/refmachine.py interprets the compiled routines of
the Lattigo fork (ring / rlwe / ckks packages) and of package main from their disassembly; objects
(rings, basis extender, key switcher, evaluator) are built by the reference's own constructors
(ring.NewRing, ring.NewFastBasisExtender, rlwe.NewKeySwitcher, ckks.NewEvaluator) run the same way.
Inputs are the repo's seeded synthetic operands (optimal_conv_b200/synth.py), so only SHA-256 digests of
the outputs are stored: tests/golden/ref_eval_vectors.json, consumed by tests/test_ref_eval_vectors.py.
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, HERE)

import common  # noqa: E402
from refmachine import CKKS, RING, RLWE, GoPanic, Machine, b2f, f2b  # noqa: E402
from optimal_conv_b200 import params as PR, synth  # noqa: E402

OUT = os.path.join(HERE, "ref_eval_vectors.json")
OPERAND_CT = "go.itab.*" + CKKS + "Ciphertext," + CKKS + "Operand"
OPERAND_PT = "go.itab.*" + CKKS + "Plaintext," + CKKS + "Operand"


def ints(a):
    return [int(v) for v in a]


def monomials(m, rQ, logN, q0):
    """pl_idx[i] = NTT(X^(2^i)) at level 0 (conv.go:241-254), transformed by the reference's NTT"""
    N = 1 << logN
    out = []
    for i in range(logN):
        mono = [0] * N
        mono[1 << i] = 1
        pin, pout = m.new_poly([mono]), m.new_poly([[0] * N])
        m.call(RING + "(*Ring).NTTLvl", [rQ, 0, pin, pout])
        out.append(m.read_poly(pout)[0])
    return out


# ---------------------------------------------------------------- ring / rlwe level
def ring_cases():
    out = {}
    logN, N = 8, 256
    for name, Q, P in (("set6_a1", PR.Q_SET6[:3], PR.P_ALL[:1]), ("set7_a2", PR.Q_SET7[:4], PR.P_ALL[:2]),
                       ("set6_a5", PR.Q_SET6[:6], PR.P_ALL[:5])):
        m = Machine()
        rQ, rP = m.new_ring(N, Q), m.new_ring(N, P)
        rec = {"logN": logN, "Q": ["%x" % q for q in Q], "P": ["%x" % p for p in P], "div_round": {}, "moddown": {},
               "keyswitch": {}, "decompose": {}}
        # tables as the reference's genNTTParams builds them
        rec["tables"] = {"psi": [common.sha(m.read_slice_u64(m.rq(r + 192) + 24 * i)) for r, mods in ((rQ, Q), (rP, P)) for i in range(len(mods))],
                         "psi_inv": [common.sha(m.read_slice_u64(m.rq(r + 216) + 24 * i)) for r, mods in ((rQ, Q), (rP, P)) for i in range(len(mods))],
                         "n_inv": [m.rq(m.rq(r + 240) + 8 * i) for r, mods in ((rQ, Q), (rP, P)) for i in range(len(mods))],
                         "mred": [m.rq(m.rq(r + 96) + 8 * i) for r, mods in ((rQ, Q), (rP, P)) for i in range(len(mods))]}
        for level in range(1, len(Q)):
            a = [ints(synth.uniform_mod(10 + i, N, Q[i])) for i in range(level + 1)]
            pin, pout = m.new_poly(a), m.new_poly([[0] * N for _ in range(level + 1)])
            m.call(RING + "(*Ring).divRoundByLastModulusNTT", [rQ, level, pin, pout])
            rec["div_round"][str(level)] = common.sha(np.array(m.read_poly(pout, level), dtype=np.uint64))
        be = m.call(RING + "NewFastBasisExtender", [rQ, rP, 0])[2]
        for level in range(0, len(Q)):
            aQ = [ints(synth.uniform_mod(20 + i, N, Q[i])) for i in range(level + 1)]
            aP = [ints(synth.uniform_mod(30 + i, N, P[i])) for i in range(len(P))]
            pQ, pP, p2 = m.new_poly(aQ), m.new_poly(aP), m.new_poly([[0] * N for _ in range(level + 1)])
            m.call(RING + "(*FastBasisExtender).ModDownSplitNTTPQ", [be, level, pQ, pP, p2])
            rec["moddown"][str(level)] = common.sha(np.array(m.read_poly(p2), dtype=np.uint64))
        params = [logN] + m.slice_u64(Q) + m.slice_u64(P) + [f2b(3.2), rQ, rP, 0]
        ks = m.call(RLWE + "NewKeySwitcher", params + [0])[11]
        beta_full = (len(Q) + len(P) - 1) // len(P)
        swk = np.stack([np.stack([synth.uniform_limbs(7000 + 10 * d + k, list(Q) + list(P), N) for k in range(2)])
                        for d in range(beta_full)])
        key = m.new_swk(swk)
        for level in range(0, len(Q)):
            c1 = synth.uniform_limbs(41 + level, Q[:level + 1], N)
            cx = m.new_poly([ints(c1[i]) for i in range(level + 1)])
            m.wb(cx + 24, 1)
            p0, p1 = m.new_poly([[0] * N for _ in Q]), m.new_poly([[0] * N for _ in Q])
            m.call(RLWE + "(*KeySwitcher).SwitchKeysInPlace", [ks, level, cx, key, p0, p1])
            rec["keyswitch"][str(level)] = [common.sha(np.array(m.read_poly(p, level + 1), dtype=np.uint64)) for p in (p0, p1)]
        rec["interpreted_instructions"] = m.steps
        out[name] = rec
        print("ring case", name, m.steps, flush=True)
    return out


# ---------------------------------------------------------------- the conv path
def conv_case(logN, B, norm, seed, out_scale, Q2, P1, bias=True, ct_scale=PR.SCALE, pt_scale=PR.SCALE):
    """main.conv_then_pack (conv.go:522-546) + evaluator.Add(ct, pl_bn_b, ct) (eval.go:258) on the seeded workload"""
    N = 1 << logN
    t0 = time.time()
    w = synth.conv_workload(Q2, P1, logN, B, seed)
    m = Machine()
    keys = {(1 << (j + 1)) + 1: w["keys"][j] for j in w["keys"]}
    params, ev = m.new_evaluator(logN, Q2, P1, PR.SCALE, keys)
    rQ = params[8]
    c0, c1 = w["ct"][0]
    ct = m.new_ct([[ints(c0[i]) for i in range(2)], [ints(c1[i]) for i in range(2)]], ct_scale)
    pl_ker = [m.new_pt([ints(w["pt_ker"][b][i]) for i in range(2)], pt_scale) if b % norm == 0 else 0 for b in range(B)]
    idx = monomials(m, rQ, logN, Q2[0])
    pl_idx = [m.new_pt([idx[i]], 1.0) for i in range(logN)]
    args = params + ev + [ct] + m.slice_u64(pl_ker) + m.slice_u64(pl_idx) + [B, norm, 1, f2b(out_scale)] + [0]
    rec = {"logN": logN, "B": B, "norm": norm, "seed": seed, "out_scale": out_scale, "ct_scale": ct_scale, "pt_scale": pt_scale,
           "Q": ["%x" % q for q in Q2], "P": ["%x" % p for p in P1], "monomials": common.sha(np.array(idx, dtype=np.uint64))}
    try:
        res = m.call("main.conv_then_pack", args, max_steps=1 << 62)[-1]
    except RuntimeError as ex:
        if isinstance(ex.__cause__, GoPanic):
            rec["panic"] = str(ex.__cause__)
            rec["interpreted_instructions"] = m.steps
            return rec
        raise
    polys, sc = m.read_ct(res)
    rec["nobias"] = {"c0": common.sha(np.array(polys[0], dtype=np.uint64)), "c1": common.sha(np.array(polys[1], dtype=np.uint64)),
                     "scale": sc, "level": len(polys[0]) - 1}
    if bias:
        pb = m.new_pt([ints(w["bias"])], out_scale)   # eval.go:240: NewPlaintext(params, 0, out_scale); :252 panics otherwise
        m.call(CKKS + "(*evaluator).Add", [ev[1], m.sym[OPERAND_CT][0], res, m.sym[OPERAND_PT][0], pb, res], max_steps=1 << 62)
        polys, sc = m.read_ct(res)
        rec["bias"] = {"c0": common.sha(np.array(polys[0], dtype=np.uint64)), "c1": common.sha(np.array(polys[1], dtype=np.uint64)),
                       "scale": sc, "level": len(polys[0]) - 1}
    rec["interpreted_instructions"] = m.steps
    print("conv case logN=%d B=%d norm=%d: %d instructions, %.0f s" % (logN, B, norm, m.steps, time.time() - t0), flush=True)
    return rec


# ---------------------------------------------------------------- evaluator ops of the baseline path
EVALOP_CASES = [
    # name, logN, Q, P, level, rotations
    ("set7_bl_level1", 8, PR.Q_SET7[:2], PR.P_PACK_BL, 1, [1, 17, 100, 127]),      # evalConv_BN_BL_test shape (alpha = 2)
    ("set6_level2_a1", 9, PR.Q_SET6[:3], PR.P_ALL[:1], 2, [3, 255]),
    ("set6_level4_a2", 8, PR.Q_SET6[:5], PR.P_ALL[:2], 4, [5]),
]


def evalop_keys(Q, P, N, gals):
    beta_full = (len(Q) + len(P) - 1) // len(P)
    return {g: np.stack([np.stack([synth.uniform_limbs(9000 + 131 * n + 10 * d + k, list(Q) + list(P), N) for k in range(2)])
                         for d in range(beta_full)]) for n, g in enumerate(gals)}


def digest_ct(m, ct):
    polys, sc = m.read_ct(ct)
    return {"c0": common.sha(np.array(polys[0], dtype=np.uint64)), "c1": common.sha(np.array(polys[1], dtype=np.uint64)),
            "scale": sc, "level": len(polys[0]) - 1}


def evalop_case(logN, Q, P, level, rots):
    """RotateHoisted / RotateNew / MulNew / Add / Sub / Rescale as the baseline path uses them
    (conv.go:133,168-171; eval.go:123,130), run by the reference's ckks.evaluator"""
    N = 1 << logN
    m = Machine()
    gals = [pow(5, r, 2 * N) for r in rots]
    keys = evalop_keys(Q, P, N, gals + [2 * N - 1])     # the last one: conjugation (GaloisElementForRowRotation)
    beta_full = (len(Q) + len(P) - 1) // len(P)
    rlk = np.stack([np.stack([synth.uniform_limbs(8000 + 10 * d + k, list(Q) + list(P), N) for k in range(2)]) for d in range(beta_full)])
    params, ev = m.new_evaluator(logN, Q, P, PR.SCALE, keys, rlk)
    e = ev[1]
    lim = lambda seed: [ints(l) for l in synth.uniform_limbs(seed, Q[:level + 1], N)]  # noqa: E731
    ct = m.new_ct([lim(61), lim(62)], PR.SCALE)
    ct2 = m.new_ct([lim(63), lim(64)], PR.SCALE)
    pt = m.new_pt(lim(65), PR.SCALE)
    rec = {"logN": logN, "Q": ["%x" % q for q in Q], "P": ["%x" % p for p in P], "level": level, "rotations": rots,
           "galois": gals, "hoisted": {}, "rotate_new": {}}
    h = m.call(CKKS + "(*evaluator).RotateHoisted", [e, ct] + m.slice_u64(rots) + [0])[-1]
    for r in rots:
        rec["hoisted"][str(r)] = digest_ct(m, m.rq(m.maps[h][r]))
    for r in rots[:2]:
        rec["rotate_new"][str(r)] = digest_ct(m, m.call(CKKS + "(*evaluator).RotateNew", [e, ct, r, 0])[-1])
    ict, ipt = m.sym[OPERAND_CT][0], m.sym[OPERAND_PT][0]
    prod = m.call(CKKS + "(*evaluator).MulNew", [e, ict, ct, ipt, pt, 0])[-1]
    rec["mul_pt"] = digest_ct(m, prod)
    rec["add"] = digest_ct(m, m.call(CKKS + "(*evaluator).AddNew", [e, ict, ct, ict, ct2, 0])[-1])
    rec["sub"] = digest_ct(m, m.call(CKKS + "(*evaluator).SubNew", [e, ict, ct, ict, ct2, 0])[-1])
    # ct x ct: MulRelinNew (mulRelin ct branch: tensor product + relinearisation with rlk), and the squaring case
    rec["mul_relin"] = digest_ct(m, m.call(CKKS + "(*evaluator).MulRelinNew", [e, ict, ct, ict, ct2, 0])[-1])
    rec["square_relin"] = digest_ct(m, m.call(CKKS + "(*evaluator).MulRelinNew", [e, ict, ct, ict, ct, 0])[-1])
    # Add / Sub with different scales and every aliasing of the receiver (evaluateInPlace's scale matching)
    rec["addsub_scaled"] = {}
    for op in ("Add", "Sub"):
        for tag, sa, sb in (("big_small", PR.SCALE * 12345.678, PR.SCALE), ("small_big", PR.SCALE, PR.SCALE * 12345.678), ("x1.5", PR.SCALE, PR.SCALE * 1.5)):
            for alias in ("a", "b", "new"):
                A, B, Cc = m.new_ct([lim(61), lim(62)], sa), m.new_ct([lim(63), lim(64)], sb), m.new_ct([lim(66), lim(67)], 7.0)
                out = {"a": A, "b": B, "new": Cc}[alias]
                m.call(CKKS + "(*evaluator)." + op, [e, ict, A, ict, B, out])
                rec["addsub_scaled"]["%s:%s:%s" % (op, tag, alias)] = digest_ct(m, out)
    # MultByi / DivByi in place, ConjugateNew
    for name in ("MultByi", "DivByi"):
        t = m.new_ct([lim(61), lim(62)], PR.SCALE)
        m.call(CKKS + "(*evaluator)." + name, [e, t, t])
        rec[name] = digest_ct(m, t)
    rec["conjugate"] = digest_ct(m, m.call(CKKS + "(*evaluator).ConjugateNew", [e, ct, 0])[-1])
    # MulByPow2(ct, 6, ct), in place as eval.go:476 calls it.  (Out of place the fork's ring.MulByPow2Lvl reads the
    # un-MForm'ed input and returns x * 2^n * 2^-64; the reference never does that, and neither does hec_mul_by_pow2.)
    p2 = m.new_ct([lim(61), lim(62)], PR.SCALE)
    m.call(CKKS + "(*evaluator).MulByPow2", [e, p2, 6, p2])
    rec["mul_by_pow2_6"] = digest_ct(m, p2)
    # Add(ct, pt, ct) at equal scales (eval.go:130,258 guard equality before adding)
    sum_ct = m.new_ct([lim(61), lim(62)], PR.SCALE)
    m.call(CKKS + "(*evaluator).Add", [e, ict, sum_ct, ipt, pt, sum_ct])
    rec["add_pt"] = digest_ct(m, sum_ct)
    # Rescale: a product whose scale is SCALE * q_level drops one level (L:ckks/evaluator.go:1291-1325)
    big = m.new_ct([lim(61), lim(62)], PR.SCALE * float(Q[level]))
    out = m.new_ct([lim(66), lim(67)], 1.0)
    res = m.call(CKKS + "(*evaluator).Rescale", [e, big, f2b(PR.SCALE), out, 0, 0])
    rec["rescale_err"] = bool(res[-2] or res[-1])
    rec["rescale"] = digest_ct(m, out)
    if level >= 2:   # two divisions in one call: DivRoundByLastModulusManyNTT with nbRescales = 2
        big2 = m.new_ct([lim(61), lim(62)], PR.SCALE * float(Q[level]) * float(Q[level - 1]))
        out2 = m.new_ct([lim(66), lim(67)], 1.0)
        res = m.call(CKKS + "(*evaluator).Rescale", [e, big2, f2b(PR.SCALE), out2, 0, 0])
        assert not (res[-2] or res[-1])
        rec["rescale2"] = digest_ct(m, out2)
    rec["interpreted_instructions"] = m.steps
    print("evaluator-op case logN=%d level=%d: %d instructions" % (logN, level, m.steps), flush=True)
    return rec


def pre_conv_bl_case(logN=8, in_wid=4, ker_wid=3):
    """main.preConv_BL (conv.go:120-143): the k^2 hoisted rotations i*in_wid + j of the baseline convolution,
    set 7, level 1, alpha = 2 (the evalConv_BN_BL_test shape)"""
    N, Q, P, level = 1 << logN, PR.Q_SET7[:2], PR.P_PACK_BL, 1
    h = ker_wid // 2
    rots = [i * in_wid + j for i in range(-h, h + 1) for j in range(-h, h + 1)]
    gal = {r: pow(5, r % (N // 2) if r >= 0 else (r & (2 * N - 1)), 2 * N) for r in rots if r}
    m = Machine()
    keys = {g: np.stack([np.stack([synth.uniform_limbs(9500 + 17 * (r % 997) + k, list(Q) + list(P), N) for k in range(2)])])
            for r, g in gal.items()}
    params, ev = m.new_evaluator(logN, Q, P, PR.SCALE, keys)
    lim = lambda seed: [ints(l) for l in synth.uniform_limbs(seed, Q[:level + 1], N)]  # noqa: E731
    ct = m.new_ct([lim(61), lim(62)], PR.SCALE)
    res = m.call("main.preConv_BL", ev + [ct, in_wid, ker_wid, 0, 0, 0], max_steps=1 << 62)
    ptr, ln = res[-3], res[-2]
    outs = m.read_u64s(ptr, ln)
    rec = {"logN": logN, "in_wid": in_wid, "ker_wid": ker_wid, "rotations": rots, "galois": {str(r): g for r, g in gal.items()},
           "Q": ["%x" % q for q in Q], "P": ["%x" % p for p in P], "out": [digest_ct(m, o) for o in outs],
           "interpreted_instructions": m.steps}
    print("preConv_BL case: %d instructions" % m.steps, flush=True)
    return rec


# ---------------------------------------------------------------- between-layer helpers
ENCODER_ITAB = "go.itab.*github.com/dwkim606/test_lattigo/ckks.encoderComplex128,github.com/dwkim606/test_lattigo/ckks.Encoder"
HELPER_MASK_SEED = 4100


def helper_mask(Q, N, level, n):
    """the n-th plaintext the stubbed encoder hands out: seeded uniform residues (NTT domain)"""
    return synth.uniform_limbs(HELPER_MASK_SEED + n, Q[:level + 1], N)


def layer_helper_case(logN=5):
    """main.ext_ctxt (conv.go:347-371), main.ext_double_ctxt (conv.go:374-414), main.keep_ctxt (conv.go:417-431) and
    main.postConv_BL (conv.go:146-178), interpreted.  Slot encoding is outside the path (the Go host keeps it), so the
    ckks.Encoder these routines call is a stub: EncodeNTT / Encode fill the plaintext the routine allocated
    (ckks.NewPlaintext, interpreted) with seeded residues instead of the embedding of the mask; what is pinned is
    the evaluator composition around it -- MulNew, RotateNew, Add in the reference's order, the plaintext scales
    (q_level, sqrt(q_level), params.Scale()) and the closing Rescale."""
    N, Q, P, level = 1 << logN, PR.Q_SET6[:3], PR.P_ALL[:2], 2
    rots_a, rots_m, rots_r = [1, 3, 6], [2, 5], [4, 7, 9]
    allrots = sorted(set(rots_a + rots_m + rots_r))
    gal = {r: pow(5, r, 2 * N) for r in allrots}
    m = Machine()
    keys = {g: np.stack([np.stack([synth.uniform_limbs(9700 + 31 * r + 10 * d + k, list(Q) + list(P), N) for k in range(2)])
                         for d in range((len(Q) + len(P) - 1) // len(P))]) for r, g in gal.items()}
    params, ev = m.new_evaluator(logN, Q, P, PR.SCALE, keys)
    count = [0]
    scales = []

    def fake_encode(em, ntt):
        sp = em.r[4]
        pt = em.rq(sp + 8)
        poly = em.rq(em.rq(pt))
        nl = em.rq(poly + 8)
        limbs = helper_mask(Q, N, nl - 1, count[0])
        rows = em.rq(poly)
        for i in range(nl):
            em.write_u64s(em.rq(rows + 24 * i), ints(limbs[i]))
        if ntt:
            em.wb(poly + 24, 1)
        scales.append(b2f(em.rq(pt + 8)))
        count[0] += 1

    enc_t = CKKS + "(*encoderComplex128)."
    m.hook(enc_t + "EncodeNTT", lambda em: fake_encode(em, True))
    m.hook(enc_t + "Encode", lambda em: fake_encode(em, False))
    m.hook(enc_t + "ToNTT", lambda em: em.wb(em.rq(em.rq(em.rq(em.r[4] + 8))) + 24, 1))
    enc = [m.sym[ENCODER_ITAB][0], m.alloc(64)]
    lim = lambda seed: [ints(l) for l in synth.uniform_limbs(seed, Q[:level + 1], N)]  # noqa: E731

    def idx_map(rots):
        h = m.new_map(24)
        for r in rots:
            m.map_put(h, r, m.slice_u64([1] * (N // 2)))
        return h

    rec = {"logN": logN, "Q": ["%x" % q for q in Q], "P": ["%x" % p for p in P], "level": level,
           "galois": {str(r): g for r, g in gal.items()}, "mask_seed": HELPER_MASK_SEED}
    # ext_ctxt(eval, encoder, input, r_idx, params)
    ct = m.new_ct([lim(61), lim(62)], PR.SCALE)
    count[0], scales[:] = 0, []
    res = m.call("main.ext_ctxt", ev + enc + [ct, idx_map(rots_a)] + params + [0], max_steps=1 << 62)
    rec["ext_ctxt"] = {"rots": rots_a, "pt_scales": list(scales), "out": digest_ct(m, res[-1])}
    # ext_double_ctxt(eval, encoder, input, m_idx, r_idx, params)
    count[0], scales[:] = 0, []
    res = m.call("main.ext_double_ctxt", ev + enc + [ct, idx_map(rots_m), idx_map(rots_r)] + params + [0], max_steps=1 << 62)
    rec["ext_double_ctxt"] = {"m_rots": rots_m, "r_rots": rots_r, "pt_scales": list(scales), "out": digest_ct(m, res[-1])}
    # keep_ctxt(params, eval, encoder, input, idx)
    count[0], scales[:] = 0, []
    res = m.call("main.keep_ctxt", params + ev + enc + [ct] + m.slice_u64([1] * (N // 2)) + [0], max_steps=1 << 62)
    rec["keep_ctxt"] = {"pt_scales": list(scales), "out": digest_ct(m, res[-1])}
    # postConv_BL(param, encoder, evaluator, ct_in_rots, in_wid, ker_wid, rot, pad, max_ker_rs)
    in_wid, ker_wid = 2, 3
    max_batch = (N // 2) // (in_wid * in_wid)
    cts = [m.new_ct([lim(70 + 2 * t), lim(71 + 2 * t)], PR.SCALE) for t in range(ker_wid * ker_wid)]

    def fl4():   # [ker_wid][ker_wid][max_batch][max_batch]float64 of ones
        def sl(ptrs_or_vals, leaf):
            a = m.alloc(8 * max(1, len(ptrs_or_vals)) * (1 if leaf else 3))
            if leaf:
                m.write_u64s(a, ptrs_or_vals)
            else:
                for i, h in enumerate(ptrs_or_vals):
                    m.write_u64s(a + 24 * i, h)
            return [a, len(ptrs_or_vals), len(ptrs_or_vals)]
        one = f2b(1.0)
        return sl([sl([sl([sl([one] * max_batch, True) for _ in range(max_batch)], False) for _ in range(ker_wid)], False)
                   for _ in range(ker_wid)], False)

    count[0], scales[:] = 0, []
    res = m.call("main.postConv_BL", params + enc + ev + m.slice_u64(cts) + [in_wid, ker_wid, 1, 0] + fl4() + [0], max_steps=1 << 62)
    rec["post_conv_bl"] = {"in_wid": in_wid, "ker_wid": ker_wid, "pt_scales": list(scales), "out": digest_ct(m, res[-1])}
    rec["interpreted_instructions"] = m.steps
    print("layer-helper case: %d instructions" % m.steps, flush=True)
    return rec


EVALCONV = {"logN": 8, "in_wid": 8, "ker_wid": 3, "real_ib": 2, "real_ob": 2, "norm": 2, "seed": 321}
EVALCONV_PT_SEED = 5200


def stub_encoder(m, Q, N, log):
    """ckks.Encoder whose EncodeCoeffs / EncodeNTT / Encode fill the plaintext they are handed (allocated by the caller with
    ckks.NewPlaintext, interpreted) with the n-th seeded vector, and whose ToNTT only sets the flag: slot encoding and
    the float -> residue rounding are pinned elsewhere; what runs for real is everything around them.
    log: list receiving (level, scale) per call."""
    def fill(em, ntt, pt_off):
        sp = em.r[4]
        pt = em.rq(sp + pt_off)
        poly = em.rq(em.rq(pt))
        nl = em.rq(poly + 8)
        limbs = synth.uniform_limbs(EVALCONV_PT_SEED + len(log), Q[:nl], N)
        rows = em.rq(poly)
        for i in range(nl):
            em.write_u64s(em.rq(rows + 24 * i), ints(limbs[i]))
        if ntt:
            em.wb(poly + 24, 1)
        log.append((nl - 1, b2f(em.rq(pt + 8))))
    enc_t = CKKS + "(*encoderComplex128)."
    m.hook(enc_t + "EncodeNTT", lambda em: fill(em, True, 8))
    m.hook(enc_t + "Encode", lambda em: fill(em, False, 8))
    m.hook(enc_t + "EncodeCoeffs", lambda em: fill(em, False, 32))     # (values []float64, plaintext): receiver, slice (3 words), pt
    m.hook(enc_t + "ToNTT", lambda em: em.wb(em.rq(em.rq(em.rq(em.r[4] + 8))) + 24, 1))
    return [m.sym[ENCODER_ITAB][0], m.alloc(64)]


def eval_conv_bn_case():
    """main.evalConv_BN (eval.go:224-263) on a hand-built main.context (layout from the DWARF: 328 bytes; N @8, ECD_LV @16,
    pl_idx @112, params @136, encoder @240, evaluator @288, pack_evaluator @304): prep_Ker's float code (reshape_ker,
    encode_ker_final -- real), the plaintext loop and the bias plaintext through the stub encoder, conv_then_pack, the
    scale / level check and the bias Add -- real."""
    c = EVALCONV
    logN, N = c["logN"], 1 << c["logN"]
    Q2, P1 = PR.Q_SET6[:2], PR.P_PACK
    max_bat = N // (c["in_wid"] ** 2)
    w = synth.conv_workload(Q2, P1, logN, max_bat, c["seed"])
    m = Machine()
    keys = {(1 << (j + 1)) + 1: w["keys"][j] for j in w["keys"]}
    params, ev = m.new_evaluator(logN, Q2, P1, PR.SCALE, keys)
    rQ = params[8]
    idx = monomials(m, rQ, logN, Q2[0])
    pl_idx = [m.new_pt([idx[i]], 1.0) for i in range(logN)]
    log = []
    enc = stub_encoder(m, Q2, N, log)
    cont = m.alloc(328)
    m.write_u64s(cont, [logN, N, 1])
    m.write_u64s(cont + 112, m.slice_u64(pl_idx))
    m.write_u64s(cont + 136, params)
    m.write_u64s(cont + 240, enc)
    m.write_u64s(cont + 288, ev)
    m.write_u64s(cont + 304, ev)
    c0, c1 = w["ct"][0]
    ct = m.new_ct([[ints(c0[i]) for i in range(2)], [ints(c1[i]) for i in range(2)]], PR.SCALE)

    def fslice(vals):
        a = m.alloc(8 * max(1, len(vals)))
        m.write_u64s(a, [f2b(float(v)) for v in vals])
        return [a, len(vals), len(vals)]
    k2 = c["ker_wid"] ** 2
    rng = np.random.default_rng(c["seed"])
    ker_in = rng.uniform(-1, 1, k2 * c["real_ib"] * c["real_ob"])
    bn_a, bn_b = rng.uniform(0.5, 1.5, c["real_ob"]), rng.uniform(-1, 1, c["real_ob"])
    out = {}
    for name, out_scale in (("ok", PR.SCALE), ("s25", float(1 << 25))):
        log.clear()
        res = m.call("main.evalConv_BN", [cont, ct] + fslice(ker_in) + fslice(bn_a) + fslice(bn_b)
                     + [c["in_wid"], c["ker_wid"], c["real_ib"], c["real_ob"], c["norm"], f2b(out_scale), 0, 0], max_steps=1 << 62)
        out[name] = {"out_scale": out_scale, "encodes": [[lv, sc] for lv, sc in log], "out": digest_ct(m, res[-1])}
    rec = dict(c, Q=["%x" % q for q in Q2], P=["%x" % p for p in P1], pt_seed=EVALCONV_PT_SEED, max_bat=max_bat, cases=out,
               interpreted_instructions=m.steps)
    print("evalConv_BN case: %d instructions" % m.steps, flush=True)
    return rec


LAYER = {"logN": 4, "in_wid": 2, "kp_wid": 1, "ker_wid": 1, "real_ib": 4, "real_ob": 4, "norm": 1, "pow": 5.0, "alpha": 0.0, "seed": 654}
LAYER_STOC_SPECS = [(2, [0, 1, 2], 3), (2, [0, 1, 3], 2)]


def layer_operands(N):
    """seeded operands of layer_case: pack-evaluator workload (conv keys), main-evaluator keys, bootstrapper, StoC factors"""
    Q, P = PR.Q_SET6, PR.P_ALL
    keys, kconj, rlk, b = ctos_operands(N)
    beta = (len(Q) + len(P) - 1) // len(P)
    for n1, diags, _ in LAYER_STOC_SPECS:
        for d in diags:
            for r in (d % n1, (d // n1) * n1):
                if r and r not in keys:
                    keys[r] = np.stack([np.stack([synth.uniform_limbs(9000 + 131 * r + 10 * dd + k, list(Q) + list(P), N) for k in range(2)])
                                        for dd in range(beta)])
    stoc = [({d: (synth.uniform_limbs(7800 + 100 * i + d, Q[:ml + 1], N), synth.uniform_limbs(7900 + 100 * i + d, P, N)) for d in diags},
             n1, ml, float(Q[ml])) for i, (n1, diags, ml) in enumerate(LAYER_STOC_SPECS)]
    return keys, kconj, rlk, b, stoc


def layer_case():
    """main.evalConv_BNRelu_new as the shipped binary has it (it predates the *_sparse kinds: 19 arguments, one
    bootstrapper), kind "Conv", iter = 2, interpreted end to end on a hand-built main.context over the whole 28 + 5
    modulus chain at N = 2^4: evalConv_BN on the pack evaluator (one special prime), Scale *= 2^pow,
    Bootstrapper.BootstrappConv_CtoS, evalReLU + MulByPow2 on both halves, keep_ctxt (stub encoder), BootstrappConv_StoC,
    Rescale -- the composition hec_conv_bn_relu / Oracle.conv_bn_relu restate."""
    c = LAYER
    logN, N = c["logN"], 1 << c["logN"]
    Q, P, P1 = PR.Q_SET6, PR.P_ALL, PR.P_PACK
    max_bat = N // (c["in_wid"] ** 2)
    m = Machine()
    keys, kconj, rlk, b, stoc = layer_operands(N)
    gk = {pow(5, r, 2 * N): k for r, k in keys.items()}
    gk[2 * N - 1] = kconj
    params, ev = m.new_evaluator(logN, Q, P, PR.SCALE, gk, rlk)
    w = synth.conv_workload(Q, P1, logN, max_bat, c["seed"])          # the pack evaluator: the same Q chain, one special prime
    pkeys = {(1 << (j + 1)) + 1: w["keys"][j] for j in w["keys"]}
    pparams, pev = m.new_evaluator(logN, Q, P1, PR.SCALE, pkeys)
    idx = monomials(m, pparams[8], logN, Q[0])
    pl_idx = [m.new_pt([idx[i]], 1.0) for i in range(logN)]
    log = []
    enc = stub_encoder(m, Q, N, log)
    btp = build_bootstrapper(m, logN, params, ev, b, stoc)
    ext = m.new_map(24)                                               # cont.ext_idx: map[int][][]int
    rows = m.alloc(48)
    for ul in range(2):
        m.write_u64s(rows + 24 * ul, m.slice_u64([1] * (N // 2)))
    m.map_put(ext, c["in_wid"], [rows, 2, 2])
    cont = m.alloc(328)
    m.write_u64s(cont, [logN, N, 1])
    m.wq(cont + 72, ext)
    m.write_u64s(cont + 112, m.slice_u64(pl_idx))
    m.write_u64s(cont + 136, params)
    m.write_u64s(cont + 240, enc)
    m.write_u64s(cont + 288, ev)
    m.write_u64s(cont + 304, pev)
    m.wq(cont + 320, btp)
    c0, c1 = w["ct"][0]
    ct = m.new_ct([[ints(c0[i]) for i in range(2)], [ints(c1[i]) for i in range(2)]], PR.SCALE)

    def fslice(vals):
        a = m.alloc(8 * max(1, len(vals)))
        m.write_u64s(a, [f2b(float(v)) for v in vals])
        return [a, len(vals), len(vals)]
    rng = np.random.default_rng(c["seed"])
    k2 = c["ker_wid"] ** 2
    ker_in = rng.uniform(-1, 1, k2 * c["real_ib"] * c["real_ob"])
    bn_a, bn_b = rng.uniform(0.5, 1.5, c["real_ob"]), rng.uniform(-1, 1, c["real_ob"])
    kind = b"Conv"
    ks = m.alloc(8)
    for i, ch in enumerate(kind):
        m.wb(ks + i, ch)
    for n in ("main.debugCtoS", "main.debugReLU", "main.debugStoC", "main.printDebug", "main.printDebugCfs", "main.prt_mat_norm_step"):
        if n in m.sym:
            m.hook(n, lambda em: None)
    t0 = time.time()
    res = m.call("main.evalConv_BNRelu_new", [cont, ct] + fslice(ker_in) + fslice(bn_a) + fslice(bn_b)
                 + [f2b(c["alpha"]), f2b(c["pow"]), c["in_wid"], c["kp_wid"], c["ker_wid"], c["real_ib"], c["real_ob"], c["norm"], 0, 1, 2,
                    ks, len(kind), 0, 0], max_steps=1 << 62)
    rec = dict(c, Q=["%x" % q for q in Q], P=["%x" % p for p in P], pt_seed=EVALCONV_PT_SEED, max_bat=max_bat,
               encodes=[[lv, sc] for lv, sc in log], out=digest_ct(m, res[-1]), interpreted_instructions=m.steps)
    print("evalConv_BNRelu_new case: %d instructions, %.0f s" % (m.steps, time.time() - t0), flush=True)
    return rec


# ---------------------------------------------------------------- evalReLU
RELU_CASES = [("n5_alpha0", 5, 0.0, 15), ("n6_leaky0.1", 6, 0.1, 15), ("n5_level12", 5, 0.0, 12), ("n5_level8_too_low", 5, 0.0, 8)]


def relu_case(logN, alpha, level):
    """main.evalReLU (conv.go:435-480): three EvaluatePoly + AddConstNew + DropLevel + Mul + Relinearize on a
    seeded level-`level` ciphertext over the first 16 moduli of set 6 (the ReLU primes are levels 5..15), alpha = 5"""
    N = 1 << logN
    Q, P = PR.Q_SET6[:16], PR.P_ALL
    m = Machine()
    beta_full = (len(Q) + len(P) - 1) // len(P)
    rlk = np.stack([np.stack([synth.uniform_limbs(8000 + 10 * d + k, list(Q) + list(P), N) for k in range(2)]) for d in range(beta_full)])
    params, ev = m.new_evaluator(logN, Q, P, PR.SCALE, {}, rlk)
    lim = lambda seed: [ints(l) for l in synth.uniform_limbs(seed, Q[:level + 1], N)]  # noqa: E731
    ct = m.new_ct([lim(61), lim(62)], PR.SCALE)
    rec = {"logN": logN, "alpha": alpha, "level": level, "Q": ["%x" % q for q in Q], "P": ["%x" % p for p in P]}
    try:
        res = m.call("main.evalReLU", params + ev + [ct, f2b(alpha), 0], max_steps=1 << 62)[-1]
        rec["out"] = digest_ct(m, res)
    except RuntimeError as ex:
        if not isinstance(ex.__cause__, GoPanic):
            raise
        rec["panic"] = str(ex.__cause__)
    rec["interpreted_instructions"] = m.steps
    print("evalReLU case logN=%d alpha=%g level=%d: %d instructions" % (logN, alpha, level, m.steps), flush=True)
    return rec


CHEBY_CASES = [("deg63_level15", 63, 15), ("deg30_level15", 30, 15), ("deg31_level10", 31, 10), ("deg7_level15", 7, 15)]


def cheby_coeffs(deg):
    rng = np.random.default_rng(1000 + deg)
    co = [float(x) for x in rng.uniform(-1, 1, deg + 1)]
    co[2] = co[5] = 0.0
    return co


POLY_CASES = [("deg5_level15", 5, 15), ("deg12_level15", 12, 15), ("deg31_level9", 31, 9)]


def cheby_case(deg, level, logN=5, power_basis=False):
    """ckks.(*evaluator).EvaluateCheby (L:ckks/polynomial_evaluation.go) with seeded real Chebyshev coefficients
    (degree 63 is the bootstrapper's SinDeg), first 16 moduli of set 6, alpha = 5; power_basis: EvaluatePoly with the
    same dense coefficients (evalReLU only exercises odd polynomials)"""
    N = 1 << logN
    Q, P = PR.Q_SET6[:16], PR.P_ALL
    m = Machine()
    beta_full = (len(Q) + len(P) - 1) // len(P)
    rlk = np.stack([np.stack([synth.uniform_limbs(8000 + 10 * d + k, list(Q) + list(P), N) for k in range(2)]) for d in range(beta_full)])
    params, ev = m.new_evaluator(logN, Q, P, PR.SCALE, {}, rlk)
    lim = lambda seed: [ints(l) for l in synth.uniform_limbs(seed, Q[:level + 1], N)]  # noqa: E731
    ct = m.new_ct([lim(61), lim(62)], PR.SCALE)
    co = cheby_coeffs(deg)
    carr = m.alloc(16 * (deg + 1))
    for i, c in enumerate(co):
        m.write_u64s(carr + 16 * i, [f2b(c), 0])                      # complex128{re, im}
    cheb = m.alloc(72)                                                 # ChebyshevInterpolation{Poly{maxDeg, coeffs, lead}, a, b}
    m.write_u64s(cheb, [deg, carr, deg + 1, deg + 1, 1, f2b(-1.0), 0, f2b(1.0), 0])
    res = m.call(CKKS + ("(*evaluator).EvaluatePoly" if power_basis else "(*evaluator).EvaluateCheby"),
                 [ev[1], ct, cheb, f2b(PR.SCALE), 0, 0, 0], max_steps=1 << 62)      # *ChebyshevInterpolation starts with its Poly
    assert not (res[-2] or res[-1])
    rec = {"logN": logN, "degree": deg, "level": level, "Q": ["%x" % q for q in Q], "P": ["%x" % p for p in P],
           "out": digest_ct(m, res[-3]), "interpreted_instructions": m.steps}
    print("%s case degree=%d level=%d: %d instructions" % ("EvaluatePoly" if power_basis else "EvaluateCheby", deg, level, m.steps), flush=True)
    return rec


# ---------------------------------------------------------------- hoisted linear transform (bootstrapping's CtoS / StoC)
TYPE_PTR_PTDIAGMATRIX = 0x56b280   # runtime type descriptor of *ckks.PtDiagMatrix (LinearTransform's type switch, 0x5268ac)
LT_CASES = [
    # name, logN, Q, P, ct level, matrix level, N1, diagonals
    ("mixed_a2", 5, PR.Q_SET6[:4], PR.P_ALL[:2], 3, 3, 4, [0, 1, 2, 3, 5, 8, 9, 12, 15]),
    ("no_zero_diag_a2", 5, PR.Q_SET6[:4], PR.P_ALL[:2], 3, 3, 4, [1, 2, 4, 6, 11]),
    ("giant_only_a2", 5, PR.Q_SET6[:4], PR.P_ALL[:2], 3, 3, 2, [0, 4, 8, 12]),
    ("mixed_a5_matlevel4", 6, PR.Q_SET6[:6], PR.P_ALL[:5], 5, 4, 8, [0, 1, 3, 7, 8, 9, 17, 25, 31]),
    ("baby_only_a1_set7", 5, PR.Q_SET7[:3], PR.P_ALL[:1], 2, 2, 4, [1, 2, 3]),
]


def lt_operands(Q, P, N, level, mat_level, n1, diags):
    """seeded operands of one LinearTransform: rotation keys, ciphertext limbs, diagonals over Q and P"""
    beta = (len(Q) + len(P) - 1) // len(P)
    rots = sorted({d % n1 for d in diags if d % n1} | {(d // n1) * n1 for d in diags if d // n1})
    keys = {r: np.stack([np.stack([synth.uniform_limbs(9000 + 131 * r + 10 * d + k, list(Q) + list(P), N) for k in range(2)])
                         for d in range(beta)]) for r in rots}
    ct = (synth.uniform_limbs(61, Q[:level + 1], N), synth.uniform_limbs(62, Q[:level + 1], N))
    D = {d: (synth.uniform_limbs(7000 + d, Q[:mat_level + 1], N), synth.uniform_limbs(7500 + d, P, N)) for d in diags}
    return keys, ct, D


def lt_case(logN, Q, P, level, mat_level, n1, diags):
    """ckks.(*evaluator).LinearTransform(ct, *PtDiagMatrix) -> MultiplyByDiagMatrixBSGS (L:ckks/linear_transform.go)"""
    N = 1 << logN
    m = Machine()
    keys, (a0, a1), D = lt_operands(Q, P, N, level, mat_level, n1, diags)
    params, ev = m.new_evaluator(logN, Q, P, PR.SCALE, {pow(5, r, 2 * N): k for r, k in keys.items()}, None)
    ct = m.new_ct([[ints(l) for l in a0], [ints(l) for l in a1]], PR.SCALE)
    vec = m.new_map(16)                               # Vec map[int][2]*ring.Poly, Montgomery + NTT form
    for d, (dq, dp) in D.items():
        pq, pp = m.new_poly([ints(l) for l in dq]), m.new_poly([ints(l) for l in dp])
        for p in (pq, pp):
            m.wb(p + 24, 1)
            m.wb(p + 25, 1)
        m.map_put(vec, d, [pq, pp])
    mat = m.alloc(48)                                 # PtDiagMatrix{LogSlots, N1, Level, Scale, Vec, naive, isGaussian}
    m.write_u64s(mat, [logN - 1, n1, mat_level, f2b(PR.SCALE), vec, 0])
    res = m.call(CKKS + "(*evaluator).LinearTransform", [ev[1], ct, TYPE_PTR_PTDIAGMATRIX, mat, 0, 0, 0], max_steps=1 << 62)
    outs = m.read_u64s(res[-3], res[-2])
    rec = {"logN": logN, "Q": ["%x" % q for q in Q], "P": ["%x" % p for p in P], "level": level, "mat_level": mat_level,
           "n1": n1, "diags": diags, "out": digest_ct(m, outs[0]), "interpreted_instructions": m.steps}
    print("LinearTransform case %s N1=%d: %d instructions" % (diags, n1, m.steps), flush=True)
    return rec


# ---------------------------------------------------------------- CoeffsToSlots / SlotsToCoeffs (the two halves of the split bootstrapping)
DFT_MATS = [(4, [0, 1, 2, 3, 5, 8, 9, 12, 15], 5), (2, [0, 1, 4, 6, 11], 4)]   # (N1, diagonals, level) of the factor matrices
DFT_Q, DFT_P, DFT_LOGN, DFT_LEVEL = PR.Q_SET6[:6], PR.P_ALL[:2], 5, 5


def dft_operands(N):
    Q, P = DFT_Q, DFT_P
    beta = (len(Q) + len(P) - 1) // len(P)
    rots = set()
    for n1, diags, _ in DFT_MATS:
        rots |= {d % n1 for d in diags if d % n1} | {(d // n1) * n1 for d in diags if d // n1}
    key = lambda s: np.stack([np.stack([synth.uniform_limbs(s + 10 * d + k, list(Q) + list(P), N) for k in range(2)]) for d in range(beta)])  # noqa: E731
    keys = {r: key(9000 + 131 * r) for r in sorted(rots)}
    kconj = key(9900)
    mats = []
    for mi, (n1, diags, ml) in enumerate(DFT_MATS):
        D = {d: (synth.uniform_limbs(7000 + 100 * mi + d, Q[:ml + 1], N), synth.uniform_limbs(7500 + 100 * mi + d, P, N)) for d in diags}
        mats.append((D, n1, ml, float(Q[ml])))     # encoded at the scale of the level the matrix consumes
    cts = [(synth.uniform_limbs(61 + 2 * t, Q[:DFT_LEVEL + 1], N), synth.uniform_limbs(62 + 2 * t, Q[:DFT_LEVEL + 1], N)) for t in range(2)]
    return keys, kconj, mats, cts


def dft_case():
    """ckks.CoeffsToSlots(vec, pDFTInv, eval) and ckks.SlotsToCoeffs(ct0, ct1, pDFT, eval) (L:ckks/bootstrap.go) with two
    seeded factor matrices, full packing"""
    logN = DFT_LOGN
    N = 1 << logN
    m = Machine()
    keys, kconj, mats, cts = dft_operands(N)
    gk = {pow(5, r, 2 * N): k for r, k in keys.items()}
    gk[2 * N - 1] = kconj
    params, ev = m.new_evaluator(logN, DFT_Q, DFT_P, PR.SCALE, gk, None)
    ptrs = []
    for D, n1, ml, ms in mats:
        vec = m.new_map(16)
        for d, (dq, dp) in D.items():
            pq, pp = m.new_poly([ints(l) for l in dq]), m.new_poly([ints(l) for l in dp])
            for p in (pq, pp):
                m.wb(p + 24, 1)
                m.wb(p + 25, 1)
            m.map_put(vec, d, [pq, pp])
        mat = m.alloc(48)
        m.write_u64s(mat, [logN - 1, n1, ml, f2b(ms), vec, 0])
        ptrs.append(mat)
    mk = lambda t: m.new_ct([[ints(l) for l in cts[t][0]], [ints(l) for l in cts[t][1]]], PR.SCALE)  # noqa: E731
    res = m.call(CKKS + "CoeffsToSlots", [mk(0)] + m.slice_u64(ptrs) + ev + [0, 0], max_steps=1 << 62)
    rec = {"logN": logN, "Q": ["%x" % q for q in DFT_Q], "P": ["%x" % p for p in DFT_P],
           "coeffs_to_slots": [digest_ct(m, res[-2]), digest_ct(m, res[-1])]}
    res = m.call(CKKS + "SlotsToCoeffs", [mk(0), mk(1)] + m.slice_u64(ptrs) + ev + [0], max_steps=1 << 62)
    rec["slots_to_coeffs"] = digest_ct(m, res[-1])
    res = m.call(CKKS + "SlotsToCoeffs", [mk(1), 0] + m.slice_u64(ptrs) + ev + [0], max_steps=1 << 62)
    rec["slots_to_coeffs_real_only"] = digest_ct(m, res[-1])
    rec["interpreted_instructions"] = m.steps
    print("CoeffsToSlots / SlotsToCoeffs case: %d instructions" % m.steps, flush=True)
    return rec


# ---------------------------------------------------------------- BootstrappConv_CtoS (first half of the split bootstrapping)
CTOS_SPECS, CTOS_FIELDS, BTP_STOC_SPECS = synth.CTOS_SPECS, synth.CTOS_FIELDS, synth.BTP_STOC_SPECS
CTOS_LOGN = 4
ctos_operands, btp_stoc_mats = synth.ctos_operands, synth.btp_stoc_mats   # seeded operands live in the package (bench.py uses them too)


def build_bootstrapper(m, logN, params, ev, b, stoc_mats=None):
    """a ckks.Bootstrapper laid out by hand (field offsets from the DWARF of the reference binary) around an evaluator the
    reference's own constructor built: the fields BootstrappConv_CtoS / _StoC / Bootstrapp read"""
    N = 1 << logN
    ptrs = []
    for D, n1, ml, ms in b["mats"]:
        vec = m.new_map(16)
        for d, (dq, dp) in D.items():
            pq, pp = m.new_poly([ints(l) for l in dq]), m.new_poly([ints(l) for l in dp])
            for p in (pq, pp):
                m.wb(p + 24, 1)
                m.wb(p + 25, 1)
            m.map_put(vec, d, [pq, pp])
        mat = m.alloc(48)
        m.write_u64s(mat, [logN - 1, n1, ml, f2b(ms), vec, 0])
        ptrs.append(mat)
    coeffs, lo, hi = b["cheby"]
    carr = m.alloc(16 * len(coeffs))
    for i, c in enumerate(coeffs):
        m.write_u64s(carr + 16 * i, [f2b(c), 0])
    cheb = m.alloc(72)
    m.write_u64s(cheb, [len(coeffs) - 1, carr, len(coeffs), len(coeffs), 1, f2b(lo), 0, f2b(hi), 0])
    btp = m.alloc(656)                                   # ckks.Bootstrapper
    m.wq(btp + 0, ev[1])                                 #   *evaluator
    m.write_u64s(btp + 360, params)                      #   params ckks.Parameters
    B = btp + 8                                          #   BootstrappingParameters
    m.write_u64s(B + 160, m.slice_u64(b["sine_qi"]) + [f2b(b["sinescale"])])   # SineEvalModuli{Qi, ScalingFactor}
    m.wq(B + 240, logN); m.wq(B + 248, logN - 1); m.wq(B + 256, f2b(PR.SCALE))
    m.wq(B + 280, b["sin_type"]); m.wq(B + 288, f2b(b["message_ratio"])); m.wq(B + 304, len(coeffs) - 1)
    m.wq(B + 312, b["sin_rescal"]); m.wq(B + 320, 0)
    m.wq(btp + 464, N // 2); m.wq(btp + 472, logN - 1)
    for off, k in ((496, "prescale"), (504, "postscale"), (512, "sinescale"), (520, "sqrt2pi"), (528, "sc_fac")):
        m.wq(btp + off, f2b(b[k]))
    m.wq(btp + 536, cheb)
    m.write_u64s(btp + 608, m.slice_u64(ptrs))           #   pDFTInv
    if stoc_mats is not None:                            #   pDFT (Bootstrapp and BootstrappConv_StoC read it)
        fptrs = []
        for D, n1, ml, ms in stoc_mats:
            vec = m.new_map(16)
            for d, (dq, dp) in D.items():
                pq, pp = m.new_poly([ints(l) for l in dq]), m.new_poly([ints(l) for l in dp])
                for p in (pq, pp):
                    m.wb(p + 24, 1)
                    m.wb(p + 25, 1)
                m.map_put(vec, d, [pq, pp])
            mat = m.alloc(48)
            m.write_u64s(mat, [logN - 1, n1, ml, f2b(ms), vec, 0])
            fptrs.append(mat)
        m.write_u64s(btp + 584, m.slice_u64(fptrs))
    return btp


def ctos_case(level_in, whole=False):
    """ckks.(*Bootstrapper).BootstrappConv_CtoS on a hand-built Bootstrapper (struct layout from the DWARF): SetScale /
    ScaleUp to the bootstrapping scale, modUp, ScaleUp, CoeffsToSlots, evaluateSine (EvaluateCheby + double angle), the
    fork's final MultByConst + Rescale.  N = 2^4, the whole 28-level chain, alpha = 5."""
    logN = CTOS_LOGN
    N = 1 << logN
    Q, P = PR.Q_SET6, PR.P_ALL
    m = Machine()
    keys, kconj, rlk, b = ctos_operands(N)
    gk = {pow(5, r, 2 * N): k for r, k in keys.items()}
    gk[2 * N - 1] = kconj
    params, ev = m.new_evaluator(logN, Q, P, PR.SCALE, gk, rlk)
    btp = build_bootstrapper(m, logN, params, ev, b, btp_stoc_mats(N) if whole else None)
    ct_scale = PR.SCALE * 2.0 ** 8
    lim = lambda seed: [ints(l) for l in synth.uniform_limbs(seed, Q[:level_in + 1], N)]  # noqa: E731
    ct = m.new_ct([lim(61), lim(62)], ct_scale)
    if whole:
        res = m.call(CKKS + "(*Bootstrapper).Bootstrapp", [btp, ct, 0], max_steps=1 << 62)
        rec = {"logN": logN, "level_in": level_in, "ct_scale": ct_scale, "out": digest_ct(m, res[-1]), "interpreted_instructions": m.steps}
        print("Bootstrapp case level_in=%d: %d instructions" % (level_in, m.steps), flush=True)
        return rec
    res = m.call(CKKS + "(*Bootstrapper).BootstrappConv_CtoS", [btp, ct, 0, 0, 0], max_steps=1 << 62)
    rec = {"logN": logN, "level_in": level_in, "ct_scale": ct_scale, "out": [digest_ct(m, res[-3]), digest_ct(m, res[-2])],
           "const_bits": res[-1], "interpreted_instructions": m.steps}
    print("BootstrappConv_CtoS case level_in=%d: %d instructions" % (level_in, m.steps), flush=True)
    return rec


# ---------------------------------------------------------------- sparse packing (LogSlots < LogN - 1): subSum and the repacking CoeffsToSlots
SPARSE_LOGN, SPARSE_LS, SPARSE_LEVEL = 5, 2, 5
SPARSE_MATS = [(2, [0, 1, 2, 3], 5), (2, [0, 1, 3], 4)]


def sparse_operands(N):
    Q, P = DFT_Q, DFT_P
    beta = (len(Q) + len(P) - 1) // len(P)
    key = lambda s: np.stack([np.stack([synth.uniform_limbs(s + 10 * d + k, list(Q) + list(P), N) for k in range(2)]) for d in range(beta)])  # noqa: E731
    logN = N.bit_length() - 1
    ls = logN - 3
    rots = {1 << i for i in range(ls, logN - 1)} | {1 << ls}
    for n1, diags, _ in SPARSE_MATS:
        rots |= {d % n1 for d in diags if d % n1} | {(d // n1) * n1 for d in diags if d // n1}
    keys = {r: key(9000 + 131 * r) for r in sorted(rots)}
    mats = [({d: (synth.uniform_limbs(7000 + 100 * mi + d, Q[:ml + 1], N), synth.uniform_limbs(7500 + 100 * mi + d, P, N)) for d in diags},
             n1, ml, float(Q[ml])) for mi, (n1, diags, ml) in enumerate(SPARSE_MATS)]
    ct = (synth.uniform_limbs(61, Q[:SPARSE_LEVEL + 1], N), synth.uniform_limbs(62, Q[:SPARSE_LEVEL + 1], N))
    return keys, key(9900), mats, ct, ls


def sparse_case():
    """Bootstrapper.subSum and ckks.CoeffsToSlots with LogSlots = LogN - 3 (the evaluator's scratch ciphertext at degree 1,
    as it is after any scale-matched Add)"""
    logN = SPARSE_LOGN
    N = 1 << logN
    m = Machine()
    keys, kconj, mats, (a0, a1), ls = sparse_operands(N)
    gk = {pow(5, r, 2 * N): k for r, k in keys.items()}
    gk[2 * N - 1] = kconj
    params, ev = m.new_evaluator(logN, DFT_Q, DFT_P, PR.SCALE, gk, None, log_slots=ls)
    inner = m.rq(m.rq(m.rq(ev[1] + 8) + 24))              # evaluator.evaluatorBuffers.ctxpool -> rlwe.Ciphertext
    m.wq(inner + 8, 2)                                    # len(Value) = 2
    mk = lambda: m.new_ct([[ints(l) for l in a0], [ints(l) for l in a1]], PR.SCALE)  # noqa: E731
    btp = m.alloc(656)
    m.wq(btp, ev[1])
    m.write_u64s(btp + 360, params)
    rec = {"logN": logN, "log_slots": ls, "sub_sum": digest_ct(m, m.call(CKKS + "(*Bootstrapper).subSum", [btp, mk(), 0])[-1])}
    ptrs = []
    for D, n1, ml, ms in mats:
        vec = m.new_map(16)
        for d, (dq, dp) in D.items():
            pq, pp = m.new_poly([ints(l) for l in dq]), m.new_poly([ints(l) for l in dp])
            for p in (pq, pp):
                m.wb(p + 24, 1)
                m.wb(p + 25, 1)
            m.map_put(vec, d, [pq, pp])
        mat = m.alloc(48)
        m.write_u64s(mat, [ls, n1, ml, f2b(ms), vec, 0])
        ptrs.append(mat)
    res = m.call(CKKS + "CoeffsToSlots", [mk()] + m.slice_u64(ptrs) + ev + [0, 0], max_steps=1 << 62)
    assert res[-1] == 0                                   # ct1 == nil
    rec["coeffs_to_slots"] = digest_ct(m, res[-2])
    rec["interpreted_instructions"] = m.steps
    print("sparse packing case: %d instructions" % m.steps, flush=True)
    return rec


SMALL_CONV = [
    # name, logN, B, norm, seed, out_scale, Q, P
    ("n8_B4", 8, 4, 1, 3, PR.SCALE, PR.Q_SET6[:2], PR.P_PACK),
    ("n8_B2", 8, 2, 1, 4, PR.SCALE, PR.Q_SET6[:2], PR.P_PACK),
    ("n10_B16", 10, 16, 1, 5, PR.SCALE, PR.Q_SET6[:2], PR.P_PACK),
    ("n10_B16_norm2", 10, 16, 2, 6, PR.SCALE, PR.Q_SET6[:2], PR.P_PACK),
    ("n10_B8_norm4_s25", 10, 8, 4, 7, float(1 << 25), PR.Q_SET6[:2], PR.P_PACK),
    ("n9_B8_set7", 9, 8, 1, 8, PR.SCALE, PR.Q_SET7[:2], PR.P_PACK),
    ("n8_B4_norm4_single", 8, 4, 4, 9, PR.SCALE, PR.Q_SET6[:2], PR.P_PACK),
    ("n8_B4_scale_panic", 8, 4, 4, 10, float(2 ** 80), PR.Q_SET6[:2], PR.P_PACK),
    ("n8_B256", 8, 256, 1, 11, PR.SCALE, PR.Q_SET6[:2], PR.P_PACK),
]


# ---------------------------------------------------------------- EncodeCoeffs + ToNTT (conv.go:513-514)
def encode_case(name):
    """ckks.scaleUpVecExact (the body of Encoder.EncodeCoeffs) followed by ring.NTTLvl (Encoder.ToNTT)"""
    logN, n, amp, scale, level = common.ENCODE_CASES[name]
    N = 1 << logN
    Q = PR.Q_SET6[:level + 1]
    m = Machine()
    rQ = m.new_ring(N, Q)
    vals = common.encode_values(name)
    va = m.alloc(8 * len(vals))
    m.write_u64s(va, [f2b(float(v)) for v in vals])
    pt = m.new_poly([[7] * N for _ in Q])           # stale contents: the tail must be cleared
    rows = m.rq(pt)
    m.call(CKKS + "scaleUpVecExact", [va, len(vals), len(vals), f2b(scale)] + m.slice_u64(Q) + [rows, len(Q), len(Q)])
    raw = np.array(m.read_poly(pt), dtype=np.uint64)
    m.call(RING + "(*Ring).NTTLvl", [rQ, level, pt, pt])
    ntt = np.array(m.read_poly(pt), dtype=np.uint64)
    return {"logN": logN, "n": n, "scale": scale, "level": level, "raw": common.sha(raw), "ntt": common.sha(ntt),
            "raw_head": [["%x" % int(x) for x in raw[j, :len(common.ENCODE_EDGE)]] for j in range(len(Q))],
            "interpreted_instructions": m.steps}


# ---------------------------------------------------------------- float-side host preparation (package main)
HOSTPREP = {"B": 4, "w": 8, "k": 3, "raw_w": 7, "norm": 2}


def hostprep_case():
    """main.reshape_ker / encode_ker_final (conv.go:184-237), prep_Input (main.go:1007-1041, trans = false),
    post_process (main.go:1057-1070) on small index-valued inputs: pins tests/hostprep.py"""
    import struct
    m = Machine()

    def fslice(vals):
        a = m.alloc(8 * max(1, len(vals)))
        m.write_u64s(a, [f2b(float(v)) for v in vals])
        return [a, len(vals), len(vals)]

    def rfs(hdr_words):
        return [struct.unpack("<d", struct.pack("<Q", x))[0] for x in m.read_u64s(hdr_words[0], hdr_words[1])]

    B, w, k, raw_w, norm = (HOSTPREP[x] for x in ("B", "w", "k", "raw_w", "norm"))
    ker = [float(v) for v in range(1, B * B * k * k + 1)]
    r = m.call("main.reshape_ker", fslice(ker) + [k * k, B, 0, 0, 0, 0])
    rows_ptr, nrows = r[6], r[7]
    reshaped = [rfs(m.read_u64s(rows_ptr + 24 * i, 3)) for i in range(nrows)]
    enc = []
    for i in range(B):
        r = m.call("main.encode_ker_final", [rows_ptr, nrows, nrows, 0, i, w, B, k, 0, 0, 0])
        enc.append(rfs(r[8:11]))
    N = w * w * B
    raw = [float(v) for v in range(1, raw_w * raw_w * (B // norm) + 1)]
    r = m.call("main.prep_Input", fslice(raw) + [raw_w, w, N, norm, 0, 0, 0, 0])
    prep = rfs(r[8:11])
    cfs = [float(v) for v in range(1, N + 1)]
    r = m.call("main.post_process", fslice(cfs) + [raw_w, w, 0, 0, 0])
    post = rfs(r[5:8])
    # the slot index maps of the between-layer helpers that the shipped binary has (rot_util.go): gen_keep_vec, both
    # halves; gen_comprs_fast, both halves (maps of rotation -> mask); prt_mat_one_norm (where the logits sit)
    def ri(hdr_words):
        return [x - (1 << 64) if x >> 63 else x for x in m.read_u64s(hdr_words[0], hdr_words[1])]
    idxmaps = {"vec_size": 128, "in_wid": 8, "kp_wid": 6, "keep": [], "comprs_fast": []}
    for ul in (0, 1):
        r = m.call("main.gen_keep_vec", [128, 8, 6, ul, 0, 0, 0])
        idxmaps["keep"].append(ri(r[4:7]))
        r = m.call("main.gen_comprs_fast", [128, 8, 6, 1, ul, 0, 0])
        maps = []
        for h in r[5:7]:
            maps.append({str(k - (1 << 64) if k >> 63 else k): ri(m.read_u64s(slot, 3)) for k, slot in sorted(m.maps[m._mapid(h)].items())})
        idxmaps["comprs_fast"].append(maps)
    m.hook("main.prt_vec", lambda em: None)   # printing only
    vec = [float(v) for v in range(1, 4 * 4 * 8 + 1)]
    r = m.call("main.prt_mat_one_norm", fslice(vec) + [8, 2, 2, 3, 0, 0, 0])
    idxmaps["mat_one_norm"] = rfs(r[7:10])
    return {"cfg": HOSTPREP, "reshape_ker": reshaped, "encode_ker_final": enc, "prep_input": prep, "post_process": post,
            "index_maps": idxmaps, "interpreted_instructions": m.steps}


def main():
    new = {"binary": "test_run (go1.16.6, github.com/dwkim606/test_lattigo v0.0.0-20220812213541-eb33b0555aaa)"}
    if "--full" in sys.argv:
        name = sys.argv[sys.argv.index("--full") + 1]
        cfg = [c for c in common.GOLDEN_CONFIGS if c["name"] == name][0]
        new["conv_full"] = {name: conv_case(PR.LOGN, cfg["B"], cfg["norm"], cfg["seed"], float(1 << cfg["out_log"]), common.Q2, common.P1)}
    elif "--only" in sys.argv:
        new["conv"] = {name: conv_case(logN, B, norm, seed, out_scale, Q2, P1)
                       for name, logN, B, norm, seed, out_scale, Q2, P1 in SMALL_CONV if name in sys.argv}
    else:
        ALL = ("relu", "evalops", "lt", "ctos", "btp", "ring", "conv", "encode", "hostprep", "helpers", "layer")
        groups = [g for g in ALL if "--" + g in sys.argv] or list(ALL)
        if "hostprep" in groups:
            new["hostprep"] = hostprep_case()
        if "helpers" in groups:
            new["layer_helpers"] = layer_helper_case()
            new["eval_conv_bn"] = eval_conv_bn_case()
        if "layer" in groups:
            new["layer"] = layer_case()
        if "encode" in groups:
            new["encode_coeffs"] = {name: encode_case(name) for name in common.ENCODE_CASES}
        if "relu" in groups:
            new["relu"] = {name: relu_case(logN, alpha, level) for name, logN, alpha, level in RELU_CASES}
            new["cheby"] = {name: cheby_case(deg, level) for name, deg, level in CHEBY_CASES}
            new["poly"] = {name: cheby_case(deg, level, power_basis=True) for name, deg, level in POLY_CASES}
        if "lt" in groups:
            new["linear_transform"] = {name: lt_case(*a) for name, *a in LT_CASES}
            new["dft"] = dft_case()
            new["sparse"] = sparse_case()
        if "ctos" in groups:
            new["ctos"] = {"level%d" % lv: ctos_case(lv) for lv in (1, 0, 3)}
        if "btp" in groups:
            new["bootstrapp"] = {"level%d" % lv: ctos_case(lv, whole=True) for lv in (1, 0)}
        if "evalops" in groups:
            new["evalops"] = {name: evalop_case(logN, Q, P, level, rots) for name, logN, Q, P, level, rots in EVALOP_CASES}
            new["pre_conv_bl"] = pre_conv_bl_case()
        if "ring" in groups:
            new["ring"] = ring_cases()
        if "conv" in groups:
            new["conv"] = {name: conv_case(logN, B, norm, seed, out_scale, Q2, P1)
                           for name, logN, B, norm, seed, out_scale, Q2, P1 in SMALL_CONV}
    # merge into the file as it is NOW (long --full runs may have finished meanwhile)
    data = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for k, v in new.items():
        if isinstance(v, dict) and k in ("conv_full", "conv") and isinstance(data.get(k), dict) and ("--full" in sys.argv or "--only" in sys.argv):
            data[k].update(v)
        else:
            data[k] = v
    json.dump(data, open(OUT, "w"), indent=1, sort_keys=True)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
