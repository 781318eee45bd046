"""GPU parity (run with -m gpu on a B200): every entry point of the C ABI against the CPU
oracle on the same seeded inputs, bit-exact (integer work: no tolerance), plus the
committed golden hashes and size-independent properties."""
import json
import os

import numpy as np
import pytest

import common
from optimal_conv_b200 import hostprep as hp
from optimal_conv_b200 import hec, params as PR, synth
from oracle.orc import Ct, Oracle

pytestmark = pytest.mark.gpu
N = 1 << PR.LOGN
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "conv_golden.json")))
Q2, P1 = common.Q2, common.P1


@pytest.fixture(scope="module")
def ctx():
    c = hec.Context(PR.LOGN, Q2, P1)
    yield c
    c.close()


@pytest.fixture(scope="module")
def orc():
    return Oracle(PR.LOGN, Q2, P1)


@pytest.fixture(scope="module")
def idx_np(orc):
    return orc.monomial_pts()


# ---------------------------------------------------------------- ring level
def test_ntt_every_modulus_matches_oracle_and_golden():
    allq = PR.Q_SET6 + [PR.Q_SET7[1], PR.Q_SET7[13]]
    c = hec.Context(PR.LOGN, allq, PR.P_ALL)
    o = Oracle(PR.LOGN, allq, PR.P_ALL)
    try:
        for ring, mods in ((0, allq), (1, PR.P_ALL)):
            for limb, q in enumerate(mods):
                a = synth.uniform_mod(500 + limb + 100 * ring, N, q)
                f = c.ntt(a, limb, ring)
                i = c.ntt(a, limb, ring, inverse=True)
                assert np.array_equal(f, o.ntt(a, limb, ring)), hex(q)
                assert np.array_equal(i, o.intt(a, limb, ring)), hex(q)
                g = GOLD["ntt"]["%d:%x" % (ring, q)]
                assert common.sha(f) == g["fwd"] and common.sha(i) == g["inv"]
                assert np.array_equal(c.ntt(f, limb, ring, inverse=True), a)
    finally:
        c.close()


def test_ntt_edge_inputs(ctx, orc):
    q = Q2[0]
    for a in (np.zeros(N, dtype=np.uint64), np.full(N, q - 1, dtype=np.uint64),
              np.eye(1, N, 0, dtype=np.uint64)[0], np.eye(1, N, N - 1, dtype=np.uint64)[0] * np.uint64(q - 1)):
        assert np.array_equal(ctx.ntt(a, 0), orc.ntt(a, 0))
        assert np.array_equal(ctx.ntt(a, 0, inverse=True), orc.intt(a, 0))


def test_monomial_plaintexts(ctx, idx_np):
    """pl_idx[i] = NTT(X^(2^i)) (conv.go:248-253) computed by the GPU NTT."""
    for i in (0, 3, 15):
        m = np.zeros(N, dtype=np.uint64)
        m[1 << i] = 1
        assert np.array_equal(ctx.ntt(m, 0), idx_np[i])
    assert common.sha(idx_np) == GOLD["monomials"]


# ---------------------------------------------------------------- evaluator ops
def _ct(seed, level=1):
    return synth.uniform_limbs(seed, Q2[:level + 1], N), synth.uniform_limbs(seed + 1, Q2[:level + 1], N)


def test_mul_add_sub_addpt_const(ctx, orc):
    a0, a1 = _ct(1)
    b0, b1 = _ct(3)
    pt = synth.uniform_limbs(5, Q2, N)
    A, B_ = ctx.upload_ct(a0, a1, 2.0 ** 30), ctx.upload_ct(b0, b1, 2.0 ** 30)
    P = ctx.upload_pt(pt, 2.0 ** 30)
    oa, ob = Ct(a0, a1, 2.0 ** 30), Ct(b0, b1, 2.0 ** 30)
    r = ctx.MulNew(A, P)
    ro = orc.mul_pt(oa, pt, 2.0 ** 30)
    g0, g1 = r.download()
    assert np.array_equal(g0, ro.c0) and np.array_equal(g1, ro.c1) and r.scale == ro.scale and r.level == 1
    s = ctx.AddNew(A, B_)
    so = orc.add(oa, ob)
    g0, g1 = s.download()
    assert np.array_equal(g0, so.c0) and np.array_equal(g1, so.c1)
    d = ctx.SubNew(A, B_)
    do = orc.sub(oa, ob)
    g0, g1 = d.download()
    assert np.array_equal(g0, do.c0) and np.array_equal(g1, do.c1)
    ctx.AddPt(s, P)
    spo = orc.add_pt(so, pt)
    g0, g1 = s.download()
    assert np.array_equal(g0, spo.c0) and np.array_equal(g1, spo.c1)
    for const in (2.0 ** -34, 3.0, -2.5, 12345.678):
        x = ctx.CopyNew(A)
        ctx.MultByConst(x, const)
        xo = orc.mul_const(oa, const)
        g0, g1 = x.download()
        assert np.array_equal(g0, xo.c0) and np.array_equal(g1, xo.c1) and x.scale == xo.scale, const
        x.free()


def test_set_scale_and_rescale(ctx, orc):
    a0, a1 = _ct(7)
    for scale_in, target in ((2.0 ** 60, 2.0 ** 26), (2.0 ** 60, 2.0 ** 30), (2.0 ** 55, 2.0 ** 20)):
        A = ctx.upload_ct(a0, a1, scale_in)
        ctx.SetScale(A, target)
        o = orc.set_scale(Ct(a0, a1, scale_in), target)
        g0, g1 = A.download()
        assert A.level == o.level == 0 and A.scale == o.scale == target
        assert np.array_equal(g0, o.c0) and np.array_equal(g1, o.c1)
    A = ctx.upload_ct(a0, a1, 2.0 ** 79)
    ctx.Rescale(A, 2.0 ** 30)
    o = orc.rescale(Ct(a0, a1, 2.0 ** 79), 2.0 ** 30)
    g0, g1 = A.download()
    assert A.level == 0 and A.scale == o.scale and np.array_equal(g0, o.c0) and np.array_equal(g1, o.c1)
    with pytest.raises(hec.HecError) as e:  # "cannot Rescale: input Ciphertext already at level 0"
        ctx.Rescale(A, 2.0 ** 30)
    assert e.value.code == hec.HEC_E_LEVEL


def test_rescale_three_limbs_with_61bit_and_30bit_moduli():
    Q = [PR.Q_SET6[5], PR.Q_SET6[0], PR.Q_SET6[2]]  # 30-bit, 56-bit, 61-bit (last is divided out)
    c, o = hec.Context(PR.LOGN, Q, P1), Oracle(PR.LOGN, Q, P1)
    try:
        a0, a1 = synth.uniform_limbs(21, Q, N), synth.uniform_limbs(22, Q, N)
        A = c.upload_ct(a0, a1, 2.0 ** 90)
        c.Rescale(A, 2.0 ** 30)
        r = o.rescale(Ct(a0, a1, 2.0 ** 90), 2.0 ** 30)
        g0, g1 = A.download()
        assert A.level == r.level == 1 and A.scale == r.scale
        assert np.array_equal(g0, r.c0) and np.array_equal(g1, r.c1)
    finally:
        c.close()


def test_rescale_two_divisions_in_one_call():
    """scale = 2^30 q_4 q_3 at level 4: Rescale divides twice (DivRoundByLastModulusManyNTT, nbRescales = 2);
    the oracle's result for this shape is pinned against the reference's compiled Rescale (test_ref_eval_vectors)."""
    Q = PR.Q_SET6[:5]
    c, o = hec.Context(PR.LOGN, Q, P1), Oracle(PR.LOGN, Q, P1)
    try:
        a0, a1 = synth.uniform_limbs(23, Q, N), synth.uniform_limbs(24, Q, N)
        s = PR.SCALE * float(Q[4]) * float(Q[3])
        A = c.upload_ct(a0, a1, s)
        c.Rescale(A, PR.SCALE)
        r = o.rescale(Ct(a0, a1, s), PR.SCALE)
        g0, g1 = A.download()
        assert A.level == r.level == 2 and A.scale == r.scale
        assert np.array_equal(g0, r.c0) and np.array_equal(g1, r.c1)
    finally:
        c.close()


# ---------------------------------------------------------------- key switching
def test_moddown_float_edge(ctx, orc):
    """v = floor(float64(y)/float64(P)) rounds up to 1 for y in [P-129, P-1]."""
    q0, p0 = Q2[0], P1[0]
    xq = synth.uniform_mod(11, N, q0)
    xp = synth.uniform_mod(12, N, p0)
    xp[:6] = [p0 - 1, p0 - 129, p0 - 130, p0 - 64, 0, 1]
    accQ, accP = orc.ntt(xq, 0)[None, :], orc.ntt(xp, 0, 1)[None, :]
    assert np.array_equal(ctx.moddown(accQ, accP), orc.moddown(accQ, accP))


@pytest.mark.parametrize("level", [0, 1])
def test_keyswitch_and_rotations_alpha1(ctx, orc, level):
    w = common.workload(common.GOLDEN_CONFIGS[1])
    g = (1 << 13) + 1
    ctx.upload_swk(g, w["keys"][12], 1)
    c0, c1 = w["ct"][0][0][:level + 1], w["ct"][0][1][:level + 1]
    d0, d1 = ctx.keyswitch(c1, g)
    e0, e1 = orc.keyswitch(c1, w["keys"][12])
    assert np.array_equal(d0, e0) and np.array_equal(d1, e1)
    A = ctx.upload_ct(c0, c1, PR.SCALE)
    out = ctx.CopyNew(A)
    ctx.RotateGal(A, g, out)
    r = orc.rotate_gal(Ct(c0, c1, PR.SCALE), g, w["keys"][12])
    g0, g1 = out.download()
    assert np.array_equal(g0, r.c0) and np.array_equal(g1, r.c1)
    gold = GOLD["rotate_gal_2^13+1"]["level%d" % level]
    assert (common.sha(g0), common.sha(g1)) == (gold["c0"], gold["c1"])
    ctx.RotateGal(A, g, A)  # in place, as conv.go:291 does
    g0, g1 = A.download()
    assert np.array_equal(g0, r.c0) and np.array_equal(g1, r.c1)
    with pytest.raises(hec.HecError) as e:
        ctx.RotateGal(A, 12345, A)
    assert e.value.code == hec.HEC_E_NOKEY


def test_job_table_cache_wraps_without_changing_results(orc, monkeypatch):
    """The generic kernels read their job tables from a content-addressed ring of slabs on the device.  With a ring of
    three 8 KiB slabs every operation overflows it: tables larger than a slab take blocks of their own, the ring comes
    round every few launches and retires what it held.  Rotations and a ct x ct product on ever new buffers must keep
    giving the oracle's result."""
    monkeypatch.setenv("HEC_STAGE_SLAB_KB", "8")
    monkeypatch.setenv("HEC_STAGE_RING", "3")
    c = hec.Context(PR.LOGN, Q2, P1)
    monkeypatch.delenv("HEC_STAGE_SLAB_KB")
    monkeypatch.delenv("HEC_STAGE_RING")
    try:
        w = common.workload(common.GOLDEN_CONFIGS[1])
        g = (1 << 13) + 1
        c.upload_swk(g, w["keys"][12], 1)
        c0, c1 = w["ct"][0][0], w["ct"][0][1]
        want = orc.rotate_gal(Ct(c0, c1, PR.SCALE), g, w["keys"][12])
        held = []
        for i in range(24):
            A = c.upload_ct(c0, c1, PR.SCALE)           # a new buffer every time: its tables are new to the cache
            out = c.CopyNew(A)
            c.RotateGal(A, g, out)
            g0, g1 = out.download()
            assert np.array_equal(g0, want.c0) and np.array_equal(g1, want.c1), i
            held.append(A if i % 3 else out)            # keep some buffers alive so that addresses do not simply recur
            (out if i % 3 else A).free()
        for h in held:
            h.free()
    finally:
        c.close()


def test_general_decomposition_alpha2_and_hoisted_rotations():
    """Baseline shape (main.go:416-430): level 1, two special primes -> one 2-limb digit through
    the float-assisted exact basis extension; RotateHoisted == per-rotation RotateNew."""
    Q, P = PR.Q_SET7[:2], PR.P_PACK_BL
    c, o = hec.Context(PR.LOGN, Q, P), Oracle(PR.LOGN, Q, P)
    try:
        a0, a1 = synth.uniform_limbs(31, Q, N), synth.uniform_limbs(32, Q, N)
        rots = [0, 1, -1, 64, -65]
        keys = {}
        for r in rots[1:]:
            g = o.galois_for_rotation(r)
            assert g == c.galois_for_rotation(r)
            keys[r] = np.stack([np.stack([synth.uniform_limbs(9000 + 10 * (r % 997) + k, Q + P, N) for k in range(2)])])
            c.upload_swk(g, keys[r], 1)
        A = c.upload_ct(a0, a1, PR.SCALE)
        outs = c.RotateHoisted(A, rots)
        for r in rots:
            g0, g1 = outs[r].download()
            if r == 0:
                assert np.array_equal(g0, a0) and np.array_equal(g1, a1)
                continue
            ref = o.rotate(Ct(a0, a1, PR.SCALE), r, keys[r])
            assert np.array_equal(g0, ref.c0) and np.array_equal(g1, ref.c1), r
            single = c.RotateNew(A, r)
            s0, s1 = single.download()
            assert np.array_equal(s0, ref.c0) and np.array_equal(s1, ref.c1), r
    finally:
        c.close()


def test_keyswitch_three_digits_alpha2_level4():
    """beta = 3 with a truncated single-limb last digit (copy path) at level 4, alpha = 2."""
    Q, P = PR.Q_SET6[:5], PR.P_PACK_BL
    c, o = hec.Context(PR.LOGN, Q, P), Oracle(PR.LOGN, Q, P)
    try:
        g = o.galois_for_rotation(5)
        key = np.stack([np.stack([synth.uniform_limbs(7000 + 10 * d + k, Q + P, N) for k in range(2)]) for d in range(3)])
        c.upload_swk(g, key, 4)
        for level in (4, 3, 2):
            c1 = synth.uniform_limbs(41 + level, Q[:level + 1], N)
            d0, d1 = c.keyswitch(c1, g)
            e0, e1 = o.keyswitch(c1, key)
            assert np.array_equal(d0, e0) and np.array_equal(d1, e1), level
    finally:
        c.close()


@pytest.mark.parametrize("shape", [("set6_level4_a2", PR.Q_SET6[:5], PR.P_PACK_BL), ("set6_level1_a1", PR.Q_SET6[:2], PR.P_PACK),
                                   ("set7_level5_a5", PR.Q_SET7[:6], PR.P_ALL)], ids=lambda s: s[0])
def test_mul_relin_ct_ct(shape):
    """MulRelinNew(ct, ct) and the squaring case == the oracle's composition, which is itself
    pinned against the reference's compiled mulRelin at small N (tests/test_ref_eval_vectors.py)."""
    _, Q, P = shape
    c, o = hec.Context(PR.LOGN, Q, P), Oracle(PR.LOGN, Q, P)
    try:
        level = len(Q) - 1
        rlk = np.stack([np.stack([synth.uniform_limbs(8000 + 10 * d + k, Q + P, N) for k in range(2)]) for d in range(o.beta_full)])
        a = [synth.uniform_limbs(61 + k, Q, N) for k in range(4)]
        A, B = c.upload_ct(a[0], a[1], PR.SCALE), c.upload_ct(a[2], a[3], PR.SCALE)
        with pytest.raises(hec.HecError) as e:
            c.MulRelinNew(A, B)
        assert e.value.code == hec.HEC_E_NOKEY
        c.upload_rlk(rlk, level)
        for X, Y, x, y in ((A, B, Ct(a[0], a[1], PR.SCALE), Ct(a[2], a[3], PR.SCALE)), (A, A, Ct(a[0], a[1], PR.SCALE), Ct(a[0], a[1], PR.SCALE))):
            res = c.MulRelinNew(X, Y)
            ref = o.mul_relin(x, y, rlk)
            g0, g1 = res.download()
            assert res.level == level and res.scale == ref.scale == PR.SCALE * PR.SCALE
            assert np.array_equal(g0, ref.c0) and np.array_equal(g1, ref.c1)
        # then Rescale as evalReLU does after every multiply
        c.Rescale(res, PR.SCALE)
        r2 = o.rescale(ref, PR.SCALE)
        g0, g1 = res.download()
        assert res.level == r2.level and np.array_equal(g0, r2.c0) and np.array_equal(g1, r2.c1)
    finally:
        c.close()


def test_add_sub_scale_matching_and_aliasing():
    """Add / Sub follow evaluateInPlace: level = min, scale = max, integer scale-up of the smaller-scale operand,
    for every aliasing of the receiver (oracle pinned against the reference's compiled Add/Sub for all of these)."""
    Q = PR.Q_SET6[:4]
    c, o = hec.Context(PR.LOGN, Q, PR.P_PACK_BL), Oracle(PR.LOGN, Q, PR.P_PACK_BL)
    try:
        S1, S2 = PR.SCALE * 12345.678, PR.SCALE
        for sub in (False, True):
            for sa, sb in ((S1, S2), (S2, S1), (S2, S2 * 1.5), (S2, S2)):
                for alias in ("a", "b", "new"):
                    a = Ct(synth.uniform_limbs(1, Q, N), synth.uniform_limbs(2, Q, N), sa)
                    b = Ct(synth.uniform_limbs(3, Q[:3], N), synth.uniform_limbs(4, Q[:3], N), sb)
                    A, B = c.upload_ct(a.c0, a.c1, sa), c.upload_ct(b.c0, b.c1, sb)
                    out = {"a": A, "b": B, "new": c.upload_ct(a.c0, a.c1, 7.0)}[alias]
                    (c.Sub if sub else c.Add)(A, B, out)
                    ref = o.add_matched(a, b, sub=sub)
                    g0, g1 = out.download()
                    assert out.level == ref.level == 2 and out.scale == ref.scale, (sub, sa / sb, alias)
                    assert np.array_equal(g0, ref.c0) and np.array_equal(g1, ref.c1), (sub, sa / sb, alias)
                    for x in {A, B, out}:
                        x.free()
    finally:
        c.close()


def test_evaluate_poly_and_eval_relu_level15_alpha5():
    """evalReLU (conv.go:435-480) at N = 2^16 on the first 16 moduli of set 6 (ReLU primes = levels 5..15), alpha = 5:
    libhec's host orchestration + kernels == the oracle's, whose small-N results are pinned against the reference's
    compiled main.evalReLU (tests/test_ref_eval_vectors.py).  Also the leaky variant and one EvaluatePoly alone."""
    Q, P = PR.Q_SET6[:16], PR.P_ALL
    c, o = hec.Context(PR.LOGN, Q, P), Oracle(PR.LOGN, Q, P)
    try:
        rlk = np.stack([np.stack([synth.uniform_limbs(8000 + 10 * d + k, Q + P, N) for k in range(2)]) for d in range(o.beta_full)])
        c.upload_rlk(rlk, 15)
        a = Ct(synth.uniform_limbs(61, Q, N), synth.uniform_limbs(62, Q, N), PR.SCALE)
        A = c.upload_ct(a.c0, a.c1, PR.SCALE)
        res = c.EvaluatePoly(A, Oracle.RELU_COEFFS[0], PR.SCALE, PR.SCALE)
        ref = o.evaluate_poly(a, Oracle.RELU_COEFFS[0], PR.SCALE, rlk, PR.SCALE)
        g0, g1 = res.download()
        assert res.level == ref.level == 12 and res.scale == ref.scale
        assert np.array_equal(g0, ref.c0) and np.array_equal(g1, ref.c1)
        rng = np.random.default_rng(1063)   # EvaluateCheby, degree 63 (the bootstrapper's sine degree)
        co = [float(x) for x in rng.uniform(-1, 1, 64)]
        res = c.EvaluateCheby(A, co, PR.SCALE, PR.SCALE)
        ref = o.evaluate_poly(a, co, PR.SCALE, rlk, PR.SCALE, cheby=True)
        g0, g1 = res.download()
        assert res.level == ref.level == 9 and res.scale == ref.scale
        assert np.array_equal(g0, ref.c0) and np.array_equal(g1, ref.c1)
        for alpha in (0.0, 0.1):
            res = c.evalReLU(A, alpha, PR.SCALE)
            ref = o.eval_relu(a, alpha, rlk, PR.SCALE)
            g0, g1 = res.download()
            assert res.level == ref.level == 5 and res.scale == ref.scale == PR.SCALE * PR.SCALE, alpha
            assert np.array_equal(g0, ref.c0) and np.array_equal(g1, ref.c1), alpha
        # the batched form (the loop over ct_boots[ul], eval.go:470-476): three different ciphertexts in one call
        bs = [Ct(synth.uniform_limbs(71 + 2 * t, Q, N), synth.uniform_limbs(72 + 2 * t, Q, N), PR.SCALE) for t in range(2)] + [a]
        outs = c.evalReLUMany([c.upload_ct(b.c0, b.c1, PR.SCALE) for b in bs], 0.1, PR.SCALE)
        for b, r in zip(bs, outs):
            ref = o.eval_relu(b, 0.1, rlk, PR.SCALE)
            g0, g1 = r.download()
            assert r.level == ref.level and r.scale == ref.scale
            assert np.array_equal(g0, ref.c0) and np.array_equal(g1, ref.c1)
        low = c.upload_ct(a.c0[:3], a.c1[:3], PR.SCALE)
        with pytest.raises(hec.HecError) as e:  # checkEnoughLevels
            c.evalReLU(low, 0.0, PR.SCALE)
        assert e.value.code == hec.HEC_E_LEVEL
    finally:
        c.close()


@pytest.mark.parametrize("shape", [
    ("mixed_a5_matlevel4", PR.Q_SET6[:6], PR.P_ALL, 5, 4, 8, [0, 1, 3, 7, 8, 9, 17, 25, 31]),
    ("no_zero_diag_a2", PR.Q_SET6[:4], PR.P_ALL[:2], 3, 3, 4, [1, 2, 4, 6, 11]),
    ("giant_only_a2", PR.Q_SET6[:4], PR.P_ALL[:2], 3, 3, 2, [0, 4, 8, 12]),
    ("baby_only_a1_set7", PR.Q_SET7[:3], PR.P_ALL[:1], 2, 2, 4, [1, 2, 3]),
    ("zero_diag_only", PR.Q_SET6[:3], PR.P_ALL[:1], 2, 2, 4, [0]),
], ids=lambda s: s[0])
def test_linear_transform_bsgs(shape):
    """LinearTransform(ct, PtDiagMatrix) (hoisted BSGS) at N = 2^16 == the oracle, whose small-N
    results for the same diagonal patterns are pinned against the reference's compiled MultiplyByDiagMatrixBSGS."""
    _, Q, P, level, ml, n1, diags = shape
    c, o = hec.Context(PR.LOGN, Q, P), Oracle(PR.LOGN, Q, P)
    try:
        rots = sorted({d % n1 for d in diags if d % n1} | {(d // n1) * n1 for d in diags if d // n1})
        keys = {r: np.stack([np.stack([synth.uniform_limbs(9000 + 131 * r + 10 * d + k, Q + P, N) for k in range(2)])
                             for d in range(o.beta_full)]) for r in rots}
        a = Ct(synth.uniform_limbs(61, Q[:level + 1], N), synth.uniform_limbs(62, Q[:level + 1], N), PR.SCALE)
        D = {d: (synth.uniform_limbs(7000 + d, Q[:ml + 1], N), synth.uniform_limbs(7500 + d, P, N)) for d in diags}
        A = c.upload_ct(a.c0, a.c1, PR.SCALE)
        mat = c.upload_ptdiag(PR.LOGN - 1, n1, ml, PR.SCALE, D)
        if rots:
            with pytest.raises(hec.HecError) as e:
                c.LinearTransform(A, mat)
            assert e.value.code == hec.HEC_E_NOKEY
        for r, k in keys.items():
            c.upload_swk(c.galois_for_rotation(r), k, level)
        res = c.LinearTransform(A, mat)
        ref = o.linear_transform(a, D, n1, ml, PR.SCALE, keys)
        g0, g1 = res.download()
        assert res.level == ref.level == min(level, ml) and res.scale == ref.scale == PR.SCALE * PR.SCALE
        assert np.array_equal(g0, ref.c0) and np.array_equal(g1, ref.c1)
        c.free_ptdiag(mat)
    finally:
        c.close()


def test_coeffs_to_slots_and_slots_to_coeffs():
    """CoeffsToSlots / SlotsToCoeffs (the linear halves of the split bootstrapping) with two factor matrices at
    N = 2^16, alpha = 2 == the oracle's compositions (pinned at small N against the reference's compiled code)."""
    Q, P, level = PR.Q_SET6[:6], PR.P_ALL[:2], 5
    specs = [(4, [0, 1, 2, 3, 5, 8, 9, 12, 15], 5), (2, [0, 1, 4, 6, 11], 4)]
    c, o = hec.Context(PR.LOGN, Q, P), Oracle(PR.LOGN, Q, P)
    try:
        key = lambda s: np.stack([np.stack([synth.uniform_limbs(s + 10 * d + k, Q + P, N) for k in range(2)]) for d in range(o.beta_full)])  # noqa: E731
        rots = set()
        for n1, diags, _ in specs:
            rots |= {d % n1 for d in diags if d % n1} | {(d // n1) * n1 for d in diags if d // n1}
        keys = {r: key(9000 + 131 * r) for r in sorted(rots)}
        kconj = key(9900)
        for r, k in keys.items():
            c.upload_swk(c.galois_for_rotation(r), k, level)
        c.upload_swk(2 * N - 1, kconj, level)
        mats, hm = [], []
        for mi, (n1, diags, ml) in enumerate(specs):
            D = {d: (synth.uniform_limbs(7000 + 100 * mi + d, Q[:ml + 1], N), synth.uniform_limbs(7500 + 100 * mi + d, P, N)) for d in diags}
            mats.append((D, n1, ml, float(Q[ml])))
            hm.append(c.upload_ptdiag(PR.LOGN - 1, n1, ml, float(Q[ml]), D))
        a = Ct(synth.uniform_limbs(61, Q, N), synth.uniform_limbs(62, Q, N), PR.SCALE)
        b = Ct(synth.uniform_limbs(63, Q, N), synth.uniform_limbs(64, Q, N), PR.SCALE)
        A, B = c.upload_ct(a.c0, a.c1, PR.SCALE), c.upload_ct(b.c0, b.c1, PR.SCALE)
        g0, g1 = c.CoeffsToSlots(A, hm)
        r0, r1 = o.coeffs_to_slots(a, mats, keys, kconj)
        for g, r in ((g0, r0), (g1, r1)):
            x0, x1 = g.download()
            assert g.level == r.level == 3 and g.scale == r.scale
            assert np.array_equal(x0, r.c0) and np.array_equal(x1, r.c1)
        for second, bref in ((B, b), (None, None)):
            g = c.SlotsToCoeffs(A, second, hm)
            r = o.slots_to_coeffs(a, bref, mats, keys)
            x0, x1 = g.download()
            assert g.level == r.level and g.scale == r.scale
            assert np.array_equal(x0, r.c0) and np.array_equal(x1, r.c1)
    finally:
        c.close()


def test_bootstrapp_conv_ctos_full_chain():
    """BootstrappConv_CtoS (first half of the split bootstrapping, eval.go:447-459) at N = 2^16 over the whole
    28 + 5 modulus chain of set 6: SetScale, modUp, four hoisted linear transforms, conjugation, degree-63
    Chebyshev sine + two double-angle steps, final constant == the oracle (pinned at N = 2^4 against the
    reference's compiled BootstrappConv_CtoS); also modUp alone."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_ref_eval_vectors as G
    Q, P = PR.Q_SET6, PR.P_ALL
    c, o = hec.Context(PR.LOGN, Q, P), Oracle(PR.LOGN, Q, P)
    try:
        keys, kconj, rlk, b = G.ctos_operands(N)
        for r, k in keys.items():
            c.upload_swk(c.galois_for_rotation(r), k, 27)
        c.upload_swk(2 * N - 1, kconj, 27)
        c.upload_rlk(rlk, 27)
        mats = [c.upload_ptdiag(PR.LOGN - 1, n1, ml, ms, D) for D, n1, ml, ms in b["mats"]]
        a = Ct(synth.uniform_limbs(61, Q[:2], N), synth.uniform_limbs(62, Q[:2], N), PR.SCALE * 2.0 ** 8)
        A = c.upload_ct(a.c0, a.c1, a.scale)
        low = Ct(a.c0[:1], a.c1[:1], a.scale)
        up, uref = c.ModUp(c.upload_ct(low.c0, low.c1, low.scale)), o.mod_up(low)
        x0, x1 = up.download()
        assert up.level == uref.level == 27 and np.array_equal(x0, uref.c0) and np.array_equal(x1, uref.c1)
        fwd = G.btp_stoc_mats(N)                       # the un-split Bootstrapp of the baseline network (test_BL.go:133)
        hfwd = [c.upload_ptdiag(PR.LOGN - 1, n1, ml, ms, D) for D, n1, ml, ms in fwd]
        whole, wref = c.Bootstrapp(A, b, mats, hfwd), o.bootstrapp(a, b, fwd, keys, kconj, rlk)
        x0, x1 = whole.download()
        assert whole.level == wref.level == 12 and whole.scale == wref.scale
        assert np.array_equal(x0, wref.c0) and np.array_equal(x1, wref.c1)
        g0, g1, k = c.BootstrappConv_CtoS(A, b, mats)
        r0, r1, kref = o.bootstrapp_conv_ctos(a, b, keys, kconj, rlk)
        assert k == kref
        for g, r in ((g0, r0), (g1, r1)):
            x0, x1 = g.download()
            assert g.level == r.level == 14 and g.scale == r.scale
            assert np.array_equal(x0, r.c0) and np.array_equal(x1, r.c1)
    finally:
        c.close()


def test_sparse_packing_sub_sum_coeffs_to_slots_and_ctos():
    """Sparse packing (LogSlots = LogN - 3, as the bootstrappers of the Resnet_crop_sparse kinds): subSum, the repacking
    CoeffsToSlots (one ciphertext returned) and BootstrappConv_CtoS on top of them == the oracle"""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_ref_eval_vectors as G
    Q, P = G.DFT_Q, G.DFT_P
    c, o = hec.Context(PR.LOGN, Q, P), Oracle(PR.LOGN, Q, P)
    try:
        keys, kconj, mats, (a0, a1), ls = G.sparse_operands(N)
        assert ls == PR.LOGN - 3
        for r, k in keys.items():
            c.upload_swk(c.galois_for_rotation(r), k, 5)
        c.upload_swk(2 * N - 1, kconj, 5)
        hm = [c.upload_ptdiag(ls, n1, ml, ms, D) for D, n1, ml, ms in mats]
        a = Ct(a0, a1, PR.SCALE)
        A = c.upload_ct(a0, a1, PR.SCALE)
        S = c.CopyNew(A)
        c.SubSum(S, ls)
        ref = o.sub_sum(a, ls, keys)
        x0, x1 = S.download()
        assert np.array_equal(x0, ref.c0) and np.array_equal(x1, ref.c1)
        g0, g1 = c.CoeffsToSlots(A, hm)
        r0, r1 = o.coeffs_to_slots(a, mats, keys, kconj, ls)
        x0, x1 = g0.download()
        assert g1 is None and r1 is None and g0.level == r0.level and g0.scale == r0.scale
        assert np.array_equal(x0, r0.c0) and np.array_equal(x1, r0.c1)
    finally:
        c.close()


def test_layer_pipeline_conv_ctos_relu_keep_stoc(idx_np):
    """One layer of evalConv_BNRelu_new (eval.go:360-575) end to end on the device, through two contexts like the
    reference's two evaluators: conv_then_pack on the pack evaluator (level 1 -> 0), scale relabel, BootstrappConv_CtoS
    on the main evaluator (level 0 -> 27 -> 14), evalReLU + MulByPow2 on both halves (14 -> 4), keep_ctxt (4 -> 3),
    BootstrappConv_StoC with two factor matrices (3 -> 1), Rescale -- every hand-off of level and scale -- == the
    oracle doing the same."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_ref_eval_vectors as G
    Q, P = PR.Q_SET6, PR.P_ALL
    cp, op_ = hec.Context(PR.LOGN, Q2, P1), Oracle(PR.LOGN, Q2, P1)
    c, o = hec.Context(PR.LOGN, Q, P), Oracle(PR.LOGN, Q, P)
    try:
        # --- conv on the pack evaluator
        w = common.workload({"B": 4, "seed": 91})
        Gc = common.GpuConv(cp, w, idx_np)
        conv = cp.conv_then_pack(Gc.cts[0], Gc.ker, 1, PR.SCALE, Gc.idx, Gc.bias)
        cp.sync()
        rconv = common.oracle_conv(op_, w, 1, PR.SCALE, idx_np)
        # --- operands of the main evaluator
        keys, kconj, rlk, b = G.ctos_operands(N)
        for r, k in keys.items():
            c.upload_swk(c.galois_for_rotation(r), k, 27)
        c.upload_swk(2 * N - 1, kconj, 27)
        c.upload_rlk(rlk, 27)
        pdftinv = [c.upload_ptdiag(PR.LOGN - 1, n1, ml, ms, D) for D, n1, ml, ms in b["mats"]]
        stoc_specs = [(2, [0, 1, 2], 3), (2, [0, 1, 3], 2)]
        stoc = [({d: (synth.uniform_limbs(7800 + 100 * i + d, Q[:ml + 1], N), synth.uniform_limbs(7900 + 100 * i + d, P, N)) for d in diags},
                 n1, ml, float(Q[ml])) for i, (n1, diags, ml) in enumerate(stoc_specs)]
        pdft = [c.upload_ptdiag(PR.LOGN - 1, n1, ml, ms, D) for D, n1, ml, ms in stoc]
        mask = synth.uniform_limbs(97, Q[:5], N)
        MASK = c.upload_pt(mask, 2.0 ** 20)
        pow2 = 5
        # --- device pipeline (the conv result moves to the main context as it is: level 0 is limb q0 in both chains)
        conv.set_scale(conv.scale * 2.0 ** 8)                    # ct_conv.Scale *= 2^pow (eval.go:437)
        h0, h1, _ = c.BootstrappConv_CtoS(conv, b, pdftinv)
        halves = c.evalReLUMany([h0, h1], 0.0, PR.SCALE)
        kept = []
        for h in halves:
            c.MulByPow2(h, pow2)
            kept.append(c.keep_ctxt(h, MASK, PR.SCALE))
        out = c.BootstrappConv_StoC(kept[0], kept[1], pdft)
        c.Rescale(out, PR.SCALE)
        # --- the same on the oracle
        r = Ct(rconv.c0, rconv.c1, rconv.scale * 2.0 ** 8)
        r0, r1, _ = o.bootstrapp_conv_ctos(r, b, keys, kconj, rlk)
        rk = [o.keep_ctxt(o.mul_by_pow2(o.eval_relu(x, 0.0, rlk, PR.SCALE), pow2), mask, 2.0 ** 20, PR.SCALE) for x in (r0, r1)]
        ref = o.rescale(o.slots_to_coeffs(rk[0], rk[1], stoc, keys), PR.SCALE)
        g0, g1 = out.download()
        assert out.level == ref.level and out.scale == ref.scale
        assert np.array_equal(g0, ref.c0) and np.array_equal(g1, ref.c1)
    finally:
        c.close()
        cp.close()


@pytest.mark.parametrize("kind", ["Conv", "StrConv_sparse"])
def test_conv_bn_relu_one_call_matches_oracle(idx_np, kind):
    """hec_conv_bn_relu = evalConv_BNRelu_new (eval.go:272-575) as ONE entry point over the two evaluators, against the
    oracle's composition of the pinned pieces.
    "Conv": one convolution, both halves through keep_ctxt (eval.go:531-537).
    "StrConv_sparse": two convolutions on kernels split by output parity with norm/2, the second shifted by a monomial,
    added, a closing monomial product (eval.go:347-397), both halves through ext_double_ctxt (eval.go:499-509)."""
    Q, P = PR.Q_SET6, PR.P_ALL
    cp, op_ = hec.Context(PR.LOGN, Q2, P1), Oracle(PR.LOGN, Q2, P1)
    c, o = hec.Context(PR.LOGN, Q, P), Oracle(PR.LOGN, Q, P)
    try:
        B, pow2, two = 4, 5, kind == "StrConv_sparse"
        w = common.workload({"B": B, "seed": 93})
        w2 = common.workload({"B": B, "seed": 94})
        out_scale = PR.SCALE
        Gc = common.GpuConv(cp, w, idx_np, 2 if two else 1, out_scale)
        ker2 = [cp.upload_pt(w2["pt_ker"][i], PR.SCALE) if i % 2 == 0 else None for i in range(B)]
        bias2 = cp.upload_pt(w2["bias"][None, :], out_scale)
        mono = lambda k: (orc_mono(op_, k, 0), cp.upload_pt(orc_mono(op_, k, 0), 1.0))  # noqa: E731
        keys, kconj, rlk, b = synth.ctos_operands(N)
        mrots, rrots = [1, 2], [4, 1]
        for r in set(mrots + rrots) - set(keys):
            keys[r] = np.stack([np.stack([synth.uniform_limbs(9000 + 131 * r + 10 * d + k, Q + P, N) for k in range(2)])
                                for d in range(o.beta_full)])
        for r, k in keys.items():
            c.upload_swk(c.galois_for_rotation(r), k, 27)
        c.upload_swk(2 * N - 1, kconj, 27)
        c.upload_rlk(rlk, 27)
        pdftinv = [c.upload_ptdiag(PR.LOGN - 1, n1, ml, ms, D) for D, n1, ml, ms in b["mats"]]
        stoc_specs = [(2, [0, 1, 2], 3), (2, [0, 1, 3], 2)]
        stoc = [({d: (synth.uniform_limbs(7800 + 100 * i + d, Q[:ml + 1], N), synth.uniform_limbs(7900 + 100 * i + d, P, N)) for d in diags},
                 n1, ml, float(Q[ml])) for i, (n1, diags, ml) in enumerate(stoc_specs)]
        pdft = [c.upload_ptdiag(PR.LOGN - 1, n1, ml, ms, D) for D, n1, ml, ms in stoc]
        lv = 4                                                         # level of the halves after evalReLU
        common_kw = dict(out_scale=out_scale, pow=float(pow2), alpha=0.0, iter=2, min_scale=PR.SCALE)
        if not two:
            masks = [synth.uniform_limbs(97 + ul, Q[:lv + 1], N) for ul in range(2)]
            out = c.conv_bn_relu(cp, Gc.cts[0], pt_ker=[Gc.ker], pt_bias=[Gc.bias], norm=[1], pt_idx=Gc.idx, btp=b, ctos_mats=pdftinv,
                                 stoc_mats=pdft, keep_mask=[c.upload_pt(m, float(Q[lv])) for m in masks], **common_kw)
            ref = Oracle.conv_bn_relu(op_, o, Ct(*w["ct"][0], PR.SCALE), pt_ker=[w["pt_ker"]], pt_bias=[w["bias"]], pt_scale=PR.SCALE,
                                      norm=[1], pt_idx=idx_np, pack_keys=w["keys"], btp=b, keys=keys, key_conj=kconj, rlk=rlk,
                                      stoc_mats=stoc, keep_mask=masks, keep_scale=float(Q[lv]), **common_kw)
        else:
            (sh_np, sh), (post_np, post) = mono(3), mono(7)
            sq = float(np.sqrt(np.float64(Q[lv])))
            mk = lambda seed, rots: {r: synth.uniform_limbs(seed + r, Q[:lv + 1], N) for r in rots}  # noqa: E731
            m_np, r_np = [mk(300 + 50 * ul, mrots) for ul in range(2)], [mk(400 + 50 * ul, rrots) for ul in range(2)]
            up = lambda d: {r: c.upload_pt(v, sq) for r, v in d.items()}  # noqa: E731
            out = c.conv_bn_relu(cp, Gc.cts[0], pt_ker=[Gc.ker, ker2], pt_bias=[Gc.bias, bias2], norm=[2, 2], pt_idx=Gc.idx, btp=b,
                                 ctos_mats=pdftinv, stoc_mats=pdft, m_idx=[up(d) for d in m_np], r_idx=[up(d) for d in r_np],
                                 pt_shift2=sh, pt_post=post, **common_kw)
            ref = Oracle.conv_bn_relu(op_, o, Ct(*w["ct"][0], PR.SCALE), pt_ker=[w["pt_ker"], w2["pt_ker"]], pt_bias=[w["bias"], w2["bias"]],
                                      pt_scale=PR.SCALE, norm=[2, 2], pt_idx=idx_np, pack_keys=w["keys"], btp=b, keys=keys, key_conj=kconj,
                                      rlk=rlk, stoc_mats=stoc, m_idx=m_np, r_idx=r_np, mask_scale=sq, pt_shift2=sh_np, pt_post=post_np,
                                      **common_kw)
        g0, g1 = out.download()
        assert out.level == ref.level and out.scale == ref.scale
        assert np.array_equal(g0, ref.c0) and np.array_equal(g1, ref.c1)
    finally:
        c.close()
        cp.close()


def orc_mono(o, k, level):
    """NTT of the monomial x^k at limbs 0..level (EncodeCoeffs of a unit vector + ToNTT, scale 1; eval.go:318-330, 372-397)"""
    out = np.empty((level + 1, N), dtype=np.uint64)
    for l in range(level + 1):
        m = np.zeros(N, dtype=np.uint64)
        m[k] = 1
        out[l] = o.ntt(m, l)
    return out


# ---------------------------------------------------------------- the conv path
@pytest.mark.parametrize("cfg", common.GOLDEN_CONFIGS, ids=lambda c: c["name"])
@pytest.mark.parametrize("flags", [hec.CONV_FUSED, hec.CONV_OPLEVEL], ids=["fused", "oplevel"])
def test_conv_then_pack_matches_oracle_and_golden(orc, idx_np, cfg, flags):
    c = hec.Context(PR.LOGN, Q2, P1)
    try:
        w = common.workload(cfg)
        out_scale = float(1 << cfg["out_log"])
        G = common.GpuConv(c, w, idx_np, cfg["norm"], out_scale)
        res = c.conv_then_pack(G.cts[0], G.ker, cfg["norm"], out_scale, G.idx, G.bias, flags)
        g0, g1 = res.download()
        ref = common.oracle_conv(orc, w, cfg["norm"], out_scale, idx_np)
        assert res.level == 0 and res.scale == ref.scale == out_scale
        assert np.array_equal(g0, ref.c0) and np.array_equal(g1, ref.c1)
        gold = GOLD["conv"][cfg["name"]]
        assert (common.sha(g0), common.sha(g1)) == (gold["c0"], gold["c1"])
    finally:
        c.close()


def test_conv_without_bias_and_single_channel(orc, idx_np):
    c = hec.Context(PR.LOGN, Q2, P1)
    try:
        w = common.workload({"B": 4, "seed": 55})
        G = common.GpuConv(c, w, idx_np)
        res = c.conv_then_pack(G.cts[0], G.ker, 1, PR.SCALE, G.idx, None)
        ref = common.oracle_conv(orc, w, 1, PR.SCALE, idx_np, bias=False)
        g0, g1 = res.download()
        assert np.array_equal(g0, ref.c0) and np.array_equal(g1, ref.c1)
        # norm == B: one real channel, no pack level (conv.go:286 loop body never runs)
        res = c.conv_then_pack(G.cts[0], G.ker, 4, PR.SCALE, G.idx, G.bias)
        ref = common.oracle_conv(orc, w, 4, PR.SCALE, idx_np)
        g0, g1 = res.download()
        assert np.array_equal(g0, ref.c0) and np.array_equal(g1, ref.c1)
    finally:
        c.close()


def test_conv_scale_panic_and_missing_key(idx_np):
    c = hec.Context(PR.LOGN, Q2, P1)
    try:
        w = common.workload({"B": 4, "seed": 56})
        G = common.GpuConv(c, w, idx_np)
        # out_scale that SetScale cannot reach with one rescale: the fused path refuses (DESIGN.md:
        # the reference would carry on with a level-1/level-0 mix; conv.go:541 only fires for B/norm == 1)
        with pytest.raises(hec.HecError) as e:
            c.conv_then_pack(G.cts[0], G.ker, 1, float(2 ** 80), G.idx, G.bias, hec.CONV_FUSED)
        assert e.value.code == hec.HEC_E_SCALE
        with pytest.raises(hec.HecError) as e:  # single real channel stays at level 1 -> conv.go:541 panic
            c.conv_then_pack(G.cts[0], G.ker, 4, float(2 ** 80), G.idx, G.bias, hec.CONV_OPLEVEL)
        assert e.value.code == hec.HEC_E_SCALE
        # eval.go:252-257: a bias plaintext that is not at out_scale is the reference's second panic
        wrong = c.upload_pt(w["bias"][None, :], PR.SCALE * 2)
        for flags in (hec.CONV_FUSED, hec.CONV_OPLEVEL):
            with pytest.raises(hec.HecError) as e:
                c.conv_then_pack(G.cts[0], G.ker, 1, PR.SCALE, G.idx, wrong, flags)
            assert e.value.code == hec.HEC_E_SCALE
        with pytest.raises(hec.HecError) as e:
            c.plan(G.ker, 1, PR.SCALE, PR.SCALE, G.idx, wrong, 2)
        assert e.value.code == hec.HEC_E_SCALE
        c.L.hec_swk_drop(c.h, (1 << 16) + 1)
        for flags in (hec.CONV_FUSED, hec.CONV_OPLEVEL):
            with pytest.raises(hec.HecError) as e:
                c.conv_then_pack(G.cts[0], G.ker, 1, PR.SCALE, G.idx, G.bias, flags)
            assert e.value.code == hec.HEC_E_NOKEY
    finally:
        c.close()


def test_plan_batch_device_and_host_runs(orc, idx_np):
    """hec_plan over 3 independent ciphertexts: device-resident run, repeated run, host-buffer run."""
    c = hec.Context(PR.LOGN, Q2, P1)
    try:
        cfg = {"B": 8, "seed": 77}
        w = common.workload(cfg, n_ct=3)
        G = common.GpuConv(c, w, idx_np)
        plan = c.plan(G.ker, 1, PR.SCALE, PR.SCALE, G.idx, G.bias, 3)
        refs = [common.oracle_conv(orc, w, 1, PR.SCALE, idx_np, m=m) for m in range(3)]
        for _ in range(2):
            outs = plan.run(G.cts)
            for m in range(3):
                g0, g1 = outs[m].download()
                assert np.array_equal(g0, refs[m].c0) and np.array_equal(g1, refs[m].c1), m
        in0 = np.stack([w["ct"][m][0] for m in range(3)])
        in1 = np.stack([w["ct"][m][1] for m in range(3)])
        o0, o1 = np.zeros((3, N), dtype=np.uint64), np.zeros((3, N), dtype=np.uint64)
        c.host_register(in0)  # page-locked caller memory: asynchronous copies
        plan.run_host(in0, in1, o0, o1)
        c.host_unregister(in0)
        for m in range(3):
            assert np.array_equal(o0[m], refs[m].c0[0]) and np.array_equal(o1[m], refs[m].c1[0]), m
        # pipelined submissions (two batches in flight, copies overlap kernels): same bits
        perm = [[0, 1, 2], [2, 0, 1], [1, 2, 0], [0, 2, 1]]
        ins0 = [np.stack([w["ct"][m][0] for m in pm]) for pm in perm]
        ins1 = [np.stack([w["ct"][m][1] for m in pm]) for pm in perm]
        outs0 = [np.zeros((3, N), dtype=np.uint64) for _ in perm]
        outs1 = [np.zeros((3, N), dtype=np.uint64) for _ in perm]
        plan.span_begin()
        tickets = [plan.submit_host(ins0[i], ins1[i], outs0[i], outs1[i]) for i in range(2)]
        plan.wait(tickets[0])
        tickets += [plan.submit_host(ins0[i], ins1[i], outs0[i], outs1[i]) for i in range(2, 4)]
        for t in tickets:
            plan.wait(t)
        assert plan.span_end_ms() > 0
        for i, pm in enumerate(perm):
            for j, m in enumerate(pm):
                assert np.array_equal(outs0[i][j], refs[m].c0[0]) and np.array_equal(outs1[i][j], refs[m].c1[0]), (i, j)
        # permuted inputs give permuted outputs (independent units)
        outs = plan.run([G.cts[2], G.cts[0], G.cts[1]])
        g0, _ = outs[0].download()
        assert np.array_equal(g0, refs[2].c0)
        plan.destroy()
    finally:
        c.close()


def test_conv_linearity_property_full_size(idx_np):
    """Size-independent property at the headline shape (B=16): conv(ct_a + ct_b) == conv(ct_a) + conv(ct_b)
    up to the rescale rounding, i.e. the difference decrypts to |e| <= small; on raw residues we check the
    exactly-linear part: the pack tree (Stage B) is Z_q-linear, so feeding the same Stage-A outputs twice
    gives identical bits (determinism), and conv of the zero ciphertext is the bias alone."""
    c = hec.Context(PR.LOGN, Q2, P1)
    try:
        w = common.workload({"B": 16, "seed": 91}, n_ct=2)
        G = common.GpuConv(c, w, idx_np)
        r1 = c.conv_then_pack(G.cts[0], G.ker, 1, PR.SCALE, G.idx, G.bias)
        r2 = c.conv_then_pack(G.cts[0], G.ker, 1, PR.SCALE, G.idx, G.bias)
        a0, a1 = r1.download()
        b0, b1 = r2.download()
        assert np.array_equal(a0, b0) and np.array_equal(a1, b1)
        z = np.zeros((2, N), dtype=np.uint64)
        Z = c.upload_ct(z, z, PR.SCALE)
        rz = c.conv_then_pack(Z, G.ker, 1, PR.SCALE, G.idx, G.bias)
        z0, z1 = rz.download()
        assert np.array_equal(z0[0], w["bias"]) and not z1.any()
    finally:
        c.close()


def test_semantic_encrypt_conv_decrypt_full_size(orc):
    """test.go:43-71 end to end at N = 2^16 (B=4, w=128, k=3): encrypt on the host (oracle keygen),
    evalConv_BN hot interval on the GPU, decrypt on the host, compare with the float convolution."""
    B, w_, k, norm = 4, 128, 3, 1
    raw_w = w_ - k // 2
    rng = np.random.default_rng(3)
    raw = rng.normal(size=raw_w * raw_w * B)
    ker = rng.uniform(-1, 1, size=B * B * k * k) / (k * k)
    bn_a, bn_b = rng.uniform(0.5, 1.5, size=B), rng.uniform(-1, 1, size=B)
    sQ, sP = orc.gen_secret(21, 192)
    c = hec.Context(PR.LOGN, Q2, P1)
    try:
        for j in (14, 15):
            g = (1 << (j + 1)) + 1
            c.upload_swk(g, orc.gen_rotkey(100 + j, g, sQ, sP), 0)
        m = hp.encode_coeffs(hp.prep_input(raw, raw_w, w_, N, norm), PR.SCALE, Q2)
        c0, c1 = orc.encrypt(5, m, sQ)
        kers = hp.prep_ker_coeffs(N, ker, bn_a, w_, k, B, B, norm)
        pt_ker = [c.upload_pt(np.stack([orc.ntt(l, i) for i, l in enumerate(hp.encode_coeffs(kc, PR.SCALE, Q2))]), PR.SCALE)
                  for kc in kers]
        idx_np = orc.monomial_pts()
        idx = [c.upload_pt(idx_np[i:i + 1], 1.0) for i in range(PR.LOGN)]
        bias = c.upload_pt(orc.ntt(hp.encode_coeffs(hp.bias_coeffs(N, bn_b, w_, norm), PR.SCALE, Q2[:1])[0], 0)[None, :], PR.SCALE)
        res = c.conv_then_pack(c.upload_ct(c0, c1, PR.SCALE), pt_ker, norm, PR.SCALE, idx, bias)
        g0, g1 = res.download()
        dec = hp.decode_coeffs(orc.decrypt(Ct(g0, g1, PR.SCALE), sQ), PR.SCALE, Q2)
        got = hp.post_process(dec, raw_w, w_)
        want = hp.plain_conv_same(raw, ker, bn_a, bn_b, raw_w, k, B)
        assert np.abs(got - want).max() < 5e-3, np.abs(got - want).max()
    finally:
        c.close()


# ---------------------------------------------------------------- EncodeCoeffs + ToNTT on the device (prep_Ker, conv.go:510-515)
def _enc_values(n, amp, seed):
    u = synth.splitmix64(seed, max(n, 1)).astype(np.float64)[:n] / 2.0 ** 64
    v = (2.0 * u - 1.0) * amp
    k = min(n, len(common.ENCODE_EDGE))
    v[:k] = common.ENCODE_EDGE[:k]
    return v


@pytest.mark.parametrize("n,amp,scale,level", [(N, 4.0, 2.0 ** 30, 1), (N - 1234, 1e-3, 2.0 ** 30, 1), (N, 64.0, 2.0 ** 60, 1),
                                               (0, 1.0, 2.0 ** 30, 0), (N, 1e3, 2.0 ** 40 + 12345.0, 0), (7, 0.5, 1.0, 1)])
def test_encode_coeffs_ntt_matches_oracle(ctx, orc, n, amp, scale, level):
    """hec_encode_coeffs == scaleUpVecExact + NTTLvl of the oracle (pinned against the reference's compiled code at
    small N), over the edge values (big path, 2^63 split, negatives rounding to zero, -0.0) and a short vector"""
    v = _enc_values(n, amp, 77 + n % 13 + level)
    pt = ctx.EncodeCoeffsNTT(v, level, scale)
    assert np.array_equal(ctx.download_pt(pt), orc.encode_coeffs_ntt(v, scale, level))
    pt.free()


def test_encode_coeffs_relu_moduli_and_errors():
    Q = PR.Q_SET6[:7]                                       # 56, 49, 61, 61, 42, 30, 31 bit limbs
    c, o = hec.Context(PR.LOGN, Q, PR.P_ALL), Oracle(PR.LOGN, Q, PR.P_ALL)
    try:
        v = _enc_values(N, 2.0 ** 20, 5)
        pt = c.EncodeCoeffsNTT(v, 6, 2.0 ** 30)
        assert np.array_equal(c.download_pt(pt), o.encode_coeffs_ntt(v, 2.0 ** 30, 6))
        # a plaintext made on the device behaves like an uploaded one
        c0, c1 = synth.uniform_limbs(1, Q, N), synth.uniform_limbs(2, Q, N)
        ct = c.upload_ct(c0, c1, 2.0 ** 30)
        up = c.upload_pt(o.encode_coeffs_ntt(v, 2.0 ** 30, 6), 2.0 ** 30)
        a, b = c.MulNew(ct, pt).download(), c.MulNew(ct, up).download()
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        for bad in ([np.nan], [np.inf], np.zeros(N + 1)):
            with pytest.raises(hec.HecError):
                c.EncodeCoeffsNTT(np.array(bad, dtype=np.float64), 1, 2.0 ** 30)
        with pytest.raises(hec.HecError):
            c.EncodeCoeffsNTT(v, 7, 2.0 ** 30)
    finally:
        c.close()


def test_encode_many_crosses_launch_chunks_and_pt_round_trip(ctx, orc):
    """17 short vectors (more than one k_encode_coeffs launch holds), each checked; upload_pt -> download_pt is the identity"""
    n, count = 3000, 17
    vals = np.stack([_enc_values(n, 10.0 ** (t % 5 - 2), 900 + t) for t in range(count)])
    pts = ctx.EncodeCoeffsNTTMany(vals, 0, 2.0 ** 30)
    for t in range(count):
        assert np.array_equal(ctx.download_pt(pts[t]), orc.encode_coeffs_ntt(vals[t], 2.0 ** 30, 0)), t
        pts[t].free()
    limbs = synth.uniform_limbs(31337, Q2, N)
    pt = ctx.upload_pt(limbs, PR.SCALE)
    assert np.array_equal(ctx.download_pt(pt), limbs)
    pt.free()


def test_prep_ker_on_the_device_feeds_the_conv(orc, idx_np):
    """prep_Ker's plaintext loop (conv.go:510-515) and the bias of evalConv_BN (eval.go:238-243) encoded by
    hec_encode_coeffs_many; the conv over them == the conv over host-encoded, uploaded plaintexts == the oracle's"""
    B, w_, k, norm = 16, 64, 3, 1
    rng = np.random.default_rng(11)
    ker = rng.uniform(-1, 1, size=B * B * k * k) / (k * k)
    bn_a, bn_b = rng.uniform(0.5, 1.5, size=B), rng.uniform(-1, 1, size=B)
    kers = np.stack(hp.prep_ker_coeffs(N, ker, bn_a, w_, k, B, B, norm))
    c = hec.Context(PR.LOGN, Q2, P1)
    try:
        w = synth.conv_workload(Q2, P1, PR.LOGN, B, 4242)
        G = common.GpuConv(c, w, idx_np, norm)
        pts = c.EncodeCoeffsNTTMany(kers, 1, PR.SCALE)
        assert len(pts) == B
        want = [orc.encode_coeffs_ntt(kers[i], PR.SCALE, 1) for i in range(B)]
        for i in (0, 7, B - 1):
            assert np.array_equal(c.download_pt(pts[i]), want[i])
        bias_v = hp.bias_coeffs(N, bn_b, w_, norm)
        bias = c.EncodeCoeffsNTT(bias_v, 0, PR.SCALE)
        bias_np = orc.encode_coeffs_ntt(bias_v, PR.SCALE, 0)
        res = c.conv_then_pack(G.cts[0], pts, norm, PR.SCALE, G.idx, bias)
        up = [c.upload_pt(x, PR.SCALE) for x in want]
        ref = c.conv_then_pack(G.cts[0], up, norm, PR.SCALE, G.idx, c.upload_pt(bias_np, PR.SCALE))
        (g0, g1), (r0, r1) = res.download(), ref.download()
        assert np.array_equal(g0, r0) and np.array_equal(g1, r1)
        o, _, _ = orc.conv_then_pack(Ct(w["ct"][0][0], w["ct"][0][1], PR.SCALE), np.stack(want), PR.SCALE, norm, PR.SCALE,
                                     idx_np, w["keys"], pt_bias=bias_np[0])
        assert np.array_equal(g0, o.c0) and np.array_equal(g1, o.c1) and res.scale == o.scale
    finally:
        c.close()


def test_baseline_conv_bl_matches_oracle():
    """Rotation-per-tap baseline (test_BL.go:98-111 -> eval.go:78-134) at its real shape: parameter set 7,
    level 1 (56+61 bit), two special primes, B=4 channels -> 2 ciphertext slots-halves of w=128:
    RotateHoisted(k^2 rotations) + per output rotation: k^2 x (MulNew + Add) + RotateNew + Add + bias."""
    Q, P = PR.Q_SET7[:2], PR.P_PACK_BL
    c, o = hec.Context(PR.LOGN, Q, P), Oracle(PR.LOGN, Q, P)
    try:
        w_, k = 128, 3
        max_batch = N // (2 * w_ * w_)  # = 2
        rot_step = w_ * w_
        h = k // 2
        rots = [i * w_ + j for i in range(-h, h + 1) for j in range(-h, h + 1)] + [t * rot_step for t in range(1, max_batch)]
        keys = {}
        for r in rots:
            if r == 0:
                continue
            keys[r] = np.stack([np.stack([synth.uniform_limbs(5000 + 13 * (r % 9973) + kk, Q + P, N) for kk in range(2)])])
            c.upload_swk(o.galois_for_rotation(r), keys[r], 1)
        a0, a1 = synth.uniform_limbs(61, Q, N), synth.uniform_limbs(62, Q, N)
        taps_np = [[synth.uniform_limbs(700 + 10 * i + t, Q, N) for t in range(k * k)] for i in range(max_batch)]
        bias_np = synth.uniform_limbs(99, Q, N)
        A = c.upload_ct(a0, a1, PR.SCALE)
        taps = [[c.upload_pt(p, PR.SCALE) for p in row] for row in taps_np]
        bias = c.upload_pt(bias_np, PR.SCALE * PR.SCALE)
        res = c.conv_bl(A, w_, k, rot_step, taps, bias)
        ref = o.conv_bl(Ct(a0, a1, PR.SCALE), w_, k, rot_step, taps_np, PR.SCALE, keys, bias_np)
        g0, g1 = res.download()
        assert res.level == 1 and res.scale == ref.scale == PR.SCALE * PR.SCALE
        assert np.array_equal(g0, ref.c0) and np.array_equal(g1, ref.c1)
        wrong = c.upload_pt(bias_np, PR.SCALE)  # eval.go:127-129 "Different scale between pl_bn_b and ctxt"
        with pytest.raises(hec.HecError) as e:
            c.conv_bl(A, w_, k, rot_step, taps, wrong)
        assert e.value.code == hec.HEC_E_SCALE
    finally:
        c.close()


@pytest.mark.parametrize("B", [64, 256])
def test_conv_full_size_channel_counts(orc, idx_np, B):
    """BASELINE.json configs 3 and 4 (batch 64 / 256 rows of the (B, w) table, main.go:578-579): the whole
    pack tree (6 resp. 8 levels, galEl 2^11+1 .. 2^17-1+2) against the oracle, bit for bit.  Kernel width
    only changes plaintext *contents* (SURVEY.md finding 5), so uniform plaintexts cover k = 3, 5, 7."""
    c = hec.Context(PR.LOGN, Q2, P1)
    try:
        w = common.workload({"B": B, "seed": 400 + B})
        G = common.GpuConv(c, w, idx_np)
        res = c.conv_then_pack(G.cts[0], G.ker, 1, PR.SCALE, G.idx, G.bias)
        g0, g1 = res.download()
        ref = common.oracle_conv(orc, w, 1, PR.SCALE, idx_np, nthreads=os.cpu_count() or 1)
        assert np.array_equal(g0, ref.c0) and np.array_equal(g1, ref.c1)
        if B == 64:  # the op-level replay agrees with the fused path
            res2 = c.conv_then_pack(G.cts[0], G.ker, 1, PR.SCALE, G.idx, G.bias, hec.CONV_OPLEVEL)
            h0, h1 = res2.download()
            assert np.array_equal(h0, g0) and np.array_equal(h1, g1)
    finally:
        c.close()


def test_between_layer_helpers_alpha5():
    """ext_ctxt / ext_double_ctxt / keep_ctxt (conv.go:347-431) at level 5 of parameter
    set 6 with the main evaluator's five special primes (alpha = 5: a full 5-limb digit through the
    float-assisted basis extension + a truncated 1-limb digit), then Rescale by a 30-bit ReLU prime."""
    Q, P = PR.Q_SET6[:6], PR.P_ALL
    c, o = hec.Context(PR.LOGN, Q, P), Oracle(PR.LOGN, Q, P)
    try:
        rots = [3, -64, 4096]
        keys = {}
        for r in rots:
            keys[r] = np.stack([np.stack([synth.uniform_limbs(8000 + 17 * (r % 7919) + 3 * d + kk, Q + P, N)
                                          for kk in range(2)]) for d in range(2)])
            c.upload_swk(o.galois_for_rotation(r), keys[r], 5)
        a0, a1 = synth.uniform_limbs(71, Q, N), synth.uniform_limbs(72, Q, N)
        pscale = float(Q[5])  # NewPlaintext(params, input.Level(), float64(params.Q()[input.Level()]))
        masks_np = {r: synth.uniform_limbs(900 + i, Q, N) for i, r in enumerate(rots)}
        A = c.upload_ct(a0, a1, PR.SCALE)
        masks = {r: c.upload_pt(m, pscale) for r, m in masks_np.items()}
        oct = Ct(a0, a1, PR.SCALE)
        # ext_ctxt
        res = c.ext_ctxt(A, masks, PR.SCALE)
        ref = o.ext_ctxt(oct, masks_np, pscale, keys, PR.SCALE)
        g0, g1 = res.download()
        assert res.level == ref.level == 4 and res.scale == ref.scale
        assert np.array_equal(g0, ref.c0) and np.array_equal(g1, ref.c1)
        # ext_double_ctxt: first pass without Rescale, second pass with (conv.go:374-414)
        mid = c.ext_ctxt(A, masks)
        res2 = c.ext_ctxt(mid, masks, PR.SCALE)
        omid = o.ext_ctxt(oct, masks_np, pscale, keys)
        ref2 = o.ext_ctxt(omid, masks_np, pscale, keys, PR.SCALE)
        g0, g1 = res2.download()
        assert res2.level == ref2.level and res2.scale == ref2.scale
        assert np.array_equal(g0, ref2.c0) and np.array_equal(g1, ref2.c1)
        # keep_ctxt
        res3 = c.keep_ctxt(A, masks[3], PR.SCALE)
        ref3 = o.keep_ctxt(oct, masks_np[3], pscale, PR.SCALE)
        g0, g1 = res3.download()
        assert res3.level == ref3.level == 4 and np.array_equal(g0, ref3.c0) and np.array_equal(g1, ref3.c1)
    finally:
        c.close()


def test_full_level_keyswitch_alpha5_beta6():
    """The 'full RNS level' stress shape of SURVEY.md 8d: level 27 (28 Q limbs), alpha = 5, beta = 6
    (five 5-limb digits + one 3-limb digit), 198 MiB switching key -- what `resnet` rotations see."""
    Q, P = PR.Q_SET6, PR.P_ALL
    c, o = hec.Context(PR.LOGN, Q, P), Oracle(PR.LOGN, Q, P)
    try:
        g = o.galois_for_rotation(7)
        key = np.stack([np.stack([synth.uniform_limbs(6000 + 3 * d + kk, Q + P, N) for kk in range(2)]) for d in range(6)])
        c.upload_swk(g, key, 27)
        for level in (27, 12):
            c1 = synth.uniform_limbs(81 + level, Q[:level + 1], N)
            d0, d1 = c.keyswitch(c1, g)
            e0, e1 = o.keyswitch(c1, key)
            assert np.array_equal(d0, e0) and np.array_equal(d1, e1), level
    finally:
        c.close()


def test_conv_then_pack_plan_cache_and_object_lifetimes(orc, idx_np):
    """hec_conv_then_pack keeps its plan between calls (a per-image caller of conv.go:522); freeing or replacing what a
    cached plan reads evicts it, and a caller's own plan keeps a freed monomial / bias / key alive until it goes."""
    c = hec.Context(PR.LOGN, Q2, P1)
    try:
        w = common.workload({"B": 4, "seed": 77}, n_ct=2)
        G = common.GpuConv(c, w, idx_np)
        refs = [common.oracle_conv(orc, w, 1, PR.SCALE, idx_np, m=m) for m in range(2)]

        def check(res, m):
            g0, g1 = res.download()
            assert np.array_equal(g0, refs[m].c0) and np.array_equal(g1, refs[m].c1)
            res.free()
        assert c.plan_cache_size() == 0
        check(c.conv_then_pack(G.cts[0], G.ker, 1, PR.SCALE, G.idx, G.bias), 0)
        assert c.plan_cache_size() == 1
        n0 = c.launch_count()
        check(c.conv_then_pack(G.cts[1], G.ker, 1, PR.SCALE, G.idx, G.bias), 1)   # same plan, other input
        assert c.plan_cache_size() == 1 and c.launch_count() - n0 == 3 + 5 * 2    # no plaintext-rescale launch (one ciphertext of 4 channels: the plan does not defer)
        r = c.conv_then_pack(G.cts[0], G.ker, 1, PR.SCALE, G.idx, None)            # other arguments: a second plan
        r.free()
        assert c.plan_cache_size() == 2
        # replacing a rotation key evicts every cached plan that reads it; results stay right
        g = (1 << 16) + 1
        c.upload_swk(g, w["keys"][15], 0)
        assert c.plan_cache_size() == 0
        check(c.conv_then_pack(G.cts[0], G.ker, 1, PR.SCALE, G.idx, G.bias), 0)
        # freeing a kernel plaintext evicts too; a fresh upload of the same values gives a fresh plan
        G.ker[2].free()
        assert c.plan_cache_size() == 0
        G.ker[2] = c.upload_pt(w["pt_ker"][2], PR.SCALE)
        check(c.conv_then_pack(G.cts[1], G.ker, 1, PR.SCALE, G.idx, G.bias), 1)
        # a caller's plan outlives the handles it reads
        plan = c.plan(G.ker, 1, PR.SCALE, PR.SCALE, G.idx, G.bias, 2)
        G.bias.free()
        for i in range(PR.LOGN):
            G.idx[i].free()
        c.L.hec_swk_drop(c.h, g)
        assert c.plan_cache_size() == 0
        junk = [c.upload_pt(idx_np[i:i + 1] ^ np.uint64(1), 1.0) for i in range(4)]  # would land in the freed blocks
        outs = plan.run(G.cts)
        for m in range(2):
            check(outs[m], m)
        plan.destroy()
        for j in junk:
            j.free()
    finally:
        c.close()


def test_plan_at_the_bench_shape_64_ciphertexts_16_channels(orc, idx_np):
    """the configuration bench.py times (BASELINE.json configs[1] at 64 ciphertexts per step): every one of the 64
    level-0 results == the oracle, and a second run of the same plan gives the same bits"""
    import os
    c = hec.Context(PR.LOGN, Q2, P1)
    try:
        M = 64
        w = synth.conv_workload(Q2, P1, PR.LOGN, 16, seed=2024, n_ct=M)
        G = common.GpuConv(c, w, idx_np)
        plan = c.plan(G.ker, 1, PR.SCALE, PR.SCALE, G.idx, G.bias, M)
        outs = plan.run(G.cts)
        got = [o.download() for o in outs]
        nt = max(1, os.cpu_count() or 1)
        for m in range(M):
            ref = common.oracle_conv(orc, w, 1, PR.SCALE, idx_np, m=m, nthreads=nt)
            assert np.array_equal(got[m][0], ref.c0) and np.array_equal(got[m][1], ref.c1), m
        outs = plan.run(G.cts)
        for m in (0, 31, 63):
            g0, g1 = outs[m].download()
            assert np.array_equal(g0, got[m][0]) and np.array_equal(g1, got[m][1])
        plan.destroy()
    finally:
        c.close()


def test_conv_1024_channels_of_which_64_real(orc, idx_np):
    """the packing of the ResNet's third block (in_wid 8: max_batch = 1024, norm = 16, test.go:94-95): the first pack levels
    use g = 2^7 + 1 and 2^8 + 1, whose automorphisms leave the 256-word block but not the CTA's 4096-word tile; fused path ==
    op-level replay == oracle"""
    c = hec.Context(PR.LOGN, Q2, P1)
    try:
        B, norm = 1024, 16
        w = synth.conv_workload(Q2, P1, PR.LOGN, B, seed=1024)
        ker = [c.upload_pt(w["pt_ker"][i], PR.SCALE) if i % norm == 0 else None for i in range(B)]
        idx = [c.upload_pt(idx_np[i:i + 1], 1.0) for i in range(PR.LOGN)]
        bias = c.upload_pt(w["bias"][None, :], PR.SCALE)
        for j, k in w["keys"].items():
            c.upload_swk((1 << (j + 1)) + 1, k, 0)
        ct = c.upload_ct(w["ct"][0][0], w["ct"][0][1], PR.SCALE)
        ref = common.oracle_conv(orc, w, norm, PR.SCALE, idx_np, nthreads=os.cpu_count() or 1)
        for flags in (hec.CONV_FUSED, hec.CONV_OPLEVEL):
            g0, g1 = c.conv_then_pack(ct, ker, norm, PR.SCALE, idx, bias, flags).download()
            assert np.array_equal(g0, ref.c0) and np.array_equal(g1, ref.c1), flags
    finally:
        c.close()


def test_conv_4096_slots_is_the_last_fused_packing(idx_np):
    """max_batch = 4096 (in_wid 4), 64 real channels: the widest packing the fused path takes (g = 2^5 + 1 still maps a
    4096-word tile onto itself); its operands would be 4 GiB for the oracle, so the check is fused == op-level replay (which the
    1024-channel case above and the B <= 256 cases pin against the oracle), and 8192 is refused by the fused path"""
    c = hec.Context(PR.LOGN, Q2, P1)
    try:
        B, norm = 4096, 64
        w = synth.conv_workload(Q2, P1, PR.LOGN, 64, seed=4096)          # 64 kernels, bias, one ciphertext
        keys = synth.conv_workload(Q2, P1, PR.LOGN, B, seed=4097, kernels=False)["keys"]
        ker = [c.upload_pt(w["pt_ker"][i // norm], PR.SCALE) if i % norm == 0 else None for i in range(B)]
        idx = [c.upload_pt(idx_np[i:i + 1], 1.0) for i in range(PR.LOGN)]
        bias = c.upload_pt(w["bias"][None, :], PR.SCALE)
        for j, k in keys.items():
            c.upload_swk((1 << (j + 1)) + 1, k, 0)
        ct = c.upload_ct(w["ct"][0][0], w["ct"][0][1], PR.SCALE)
        outs = [c.conv_then_pack(ct, ker, norm, PR.SCALE, idx, bias, flags).download()
                for flags in (hec.CONV_FUSED, hec.CONV_OPLEVEL)]
        assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
        with pytest.raises(hec.HecError) as e:
            c.conv_then_pack(ct, ker + [None] * B, norm, PR.SCALE, idx, bias, hec.CONV_FUSED)
        assert e.value.code == hec.HEC_E_UNSUPPORTED
    finally:
        c.close()


def test_deferred_plan_needs_monomials_and_falls_back_without_them(orc, idx_np):
    """A plan with enough work per run (batch * channels >= 64, B <= 256) defers the forward transforms (pairs (U, e),
    DESIGN.md 4.5) after checking that the pl_idx plaintexts are the monomials X^step; with any other plaintext in their
    place the products are real products and the plan keeps the transform-per-mod-down kernels; a small batch keeps them
    too.  All must equal the oracle fed with the same plaintexts."""
    c = hec.Context(PR.LOGN, Q2, P1)
    try:
        M = 8
        w = common.workload({"B": 8, "seed": 4242}, n_ct=M)
        G = common.GpuConv(c, w, idx_np, norm=1)
        refs = [common.oracle_conv(orc, w, 1, PR.SCALE, idx_np, m=m) for m in range(M)]
        plan = c.plan(G.ker, 1, PR.SCALE, PR.SCALE, G.idx, G.bias, M)
        assert plan.deferred
        outs = plan.run(G.cts)
        for m in range(M):
            g0, g1 = outs[m].download()
            assert np.array_equal(g0, refs[m].c0) and np.array_equal(g1, refs[m].c1), m
        plan.destroy()
        small = c.plan(G.ker, 1, PR.SCALE, PR.SCALE, G.idx, G.bias, 2)
        assert not small.deferred
        outs = small.run(G.cts[:2])
        for m in range(2):
            g0, g1 = outs[m].download()
            assert np.array_equal(g0, refs[m].c0) and np.array_equal(g1, refs[m].c1), m
        small.destroy()
        odd = idx_np.copy()
        odd[1] = synth.uniform_mod(777, 1 << PR.LOGN, Q2[0])       # the level with step 2 multiplies by a random plaintext
        idx2 = [c.upload_pt(odd[i:i + 1], 1.0) for i in range(PR.LOGN)]
        plan = c.plan(G.ker, 1, PR.SCALE, PR.SCALE, idx2, G.bias, M)
        assert not plan.deferred
        outs = plan.run(G.cts)
        for m in (0, M - 1):
            ref = common.oracle_conv(orc, w, 1, PR.SCALE, odd, m=m)
            g0, g1 = outs[m].download()
            assert np.array_equal(g0, ref.c0) and np.array_equal(g1, ref.c1), m
        plan.destroy()
    finally:
        c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("B,norm,M,bias", [(16, 2, 8, True), (4, 4, 64, True), (16, 16, 64, False), (64, 1, 1, True), (32, 4, 8, False)])
def test_deferred_plan_shapes(orc, idx_np, B, norm, M, bias):
    """deferred plans over the shapes the tree can take: norm > 1 (fewer levels), a single active channel (no level: stage A
    straight into the final transform, with and without bias), one ciphertext of many channels, no bias"""
    c = hec.Context(PR.LOGN, Q2, P1)
    try:
        w = common.workload({"B": B, "seed": 5000 + B + norm}, n_ct=M)
        G = common.GpuConv(c, w, idx_np, norm=norm)
        plan = c.plan(G.ker, norm, PR.SCALE, PR.SCALE, G.idx, G.bias if bias else None, M)
        assert plan.deferred
        outs = plan.run(G.cts)
        for m in sorted({0, M // 2, M - 1}):
            ref = common.oracle_conv(orc, w, norm, PR.SCALE, idx_np, m=m, bias=bias, nthreads=os.cpu_count() or 1)
            g0, g1 = outs[m].download()
            assert np.array_equal(g0, ref.c0) and np.array_equal(g1, ref.c1), m
        plan.destroy()
    finally:
        c.close()


def test_plan_batch_with_norm_and_no_bias(orc, idx_np):
    """batch of 2 ciphertexts, B = 8 with norm = 2 (4 real channels, 2 pack levels), bias omitted"""
    c = hec.Context(PR.LOGN, Q2, P1)
    try:
        w = common.workload({"B": 8, "seed": 88}, n_ct=2)
        G = common.GpuConv(c, w, idx_np, norm=2)
        plan = c.plan(G.ker, 2, PR.SCALE, PR.SCALE, G.idx, None, 2)
        outs = plan.run(G.cts)
        for m in range(2):
            ref = common.oracle_conv(orc, w, 2, PR.SCALE, idx_np, m=m, bias=False)
            g0, g1 = outs[m].download()
            assert np.array_equal(g0, ref.c0) and np.array_equal(g1, ref.c1), m
        with pytest.raises(hec.HecError) as e:  # inputs must carry the plan's scale
            bad = c.upload_ct(w["ct"][0][0], w["ct"][0][1], PR.SCALE * 2)
            plan.run([bad, G.cts[1]])
        assert e.value.code == hec.HEC_E_SCALE
        plan.destroy()
    finally:
        c.close()


def test_resnet_chain_runs_level_by_level_and_is_deterministic():
    """optimal_conv_b200/resnet.py on the device: the depth-8 variant of the reference's chain (3 + 1 + 1 + 1 + 1 bootstrapped
    layers + the final FC, test.go:96-104) on synthetic keys / masks / matrices with the real supports -- every layer
    one hec_conv_bn_relu call that hands a level-1 ciphertext at params.Scale() to the next, the final layer a level-0
    one; two runs give the same bits."""
    import hashlib
    from optimal_conv_b200 import resnet
    net = resnet.Resnet(hec, depth=8, ker_wid=3)
    try:
        image = np.random.default_rng(5).uniform(0, 1, 32 * 32 * 3)
        digs = []
        for _ in range(2):
            ct, rec = net.run(image, seed=3)
            assert [r["layer"] for r in rec] == [s["name"] for s in net.specs] and len(rec) == 8
            assert [r["level_out"] for r in rec] == [1] * 7 + [0]
            assert ct.scale == PR.SCALE
            g0, g1 = ct.download()
            digs.append(hashlib.sha256(g0.tobytes() + g1.tobytes()).hexdigest())
            ct.free()
        assert digs[0] == digs[1]
    finally:
        net.close()
