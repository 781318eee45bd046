"""Seeded differential fuzz: random shapes of the evaluator ops and of the conv path, libhec vs the oracle, bit for
bit.  Complements the fixed-shape parity tests: levels, alpha, beta, channel counts and seeds are drawn from a PRNG
(fixed seed, so failures reproduce)."""
import numpy as np
import pytest

import common
from optimal_conv_b200 import hec, params as PR, synth
from oracle.orc import Ct, Oracle

pytestmark = pytest.mark.gpu
N = 1 << PR.LOGN


def eq(res, ref):
    g0, g1 = res.download()
    return res.level == ref.level and res.scale == ref.scale and np.array_equal(g0, ref.c0) and np.array_equal(g1, ref.c1)


@pytest.mark.parametrize("trial", range(6))
def test_fuzz_evaluator_ops(trial):
    rng = np.random.default_rng(1000 + trial)
    nQ = int(rng.integers(2, 9))
    nP = int(rng.choice([1, 2, 3, 5]))
    pool = PR.Q_SET6 if trial % 2 == 0 else PR.Q_SET7
    start = int(rng.integers(0, len(pool) - nQ))
    Q, P = pool[start:start + nQ], PR.P_ALL[:nP]
    level = int(rng.integers(1, nQ))
    c, o = hec.Context(PR.LOGN, Q, P), Oracle(PR.LOGN, Q, P)
    try:
        seed = int(rng.integers(1, 1 << 30))
        a = Ct(synth.uniform_limbs(seed, Q[:level + 1], N), synth.uniform_limbs(seed + 1, Q[:level + 1], N), PR.SCALE)
        b = Ct(synth.uniform_limbs(seed + 2, Q[:level + 1], N), synth.uniform_limbs(seed + 3, Q[:level + 1], N), PR.SCALE * 3.25)
        A, B = c.upload_ct(a.c0, a.c1, a.scale), c.upload_ct(b.c0, b.c1, b.scale)
        key = lambda s: np.stack([np.stack([synth.uniform_limbs(s + 10 * d + k, Q + P, N) for k in range(2)])  # noqa: E731
                                  for d in range(o.beta_full)])
        # rotation by a random step (RotateNew) and a Galois element of the pack tree (RotateGal)
        r = int(rng.integers(1, N // 2))
        kr = key(seed + 100)
        c.upload_swk(c.galois_for_rotation(r), kr, level)
        assert eq(c.RotateNew(A, r), o.rotate(a, r, kr)), ("rotate", Q, P, level, r)
        g = (1 << int(rng.integers(9, 17))) + 1
        kg = key(seed + 200)
        c.upload_swk(g, kg, level)
        out = c.CopyNew(A)
        c.RotateGal(A, g, out)
        assert eq(out, o.rotate_gal(a, g, kg)), ("rotate_gal", Q, P, level, g)
        # conjugation and the products with +-i
        kc = key(seed + 250)
        c.upload_swk(2 * N - 1, kc, level)
        out = c.CopyNew(A)
        c.Conjugate(A, out)
        assert eq(out, o.conjugate(a, kc)), ("conjugate", Q, P, level)
        c.MultByi(out)
        iref = o.mult_by_i(o.conjugate(a, kc))
        assert eq(out, iref), ("mult_by_i", Q, P, level)
        c.DivByi(out)
        c.DivByi(out)
        assert eq(out, o.mult_by_i(o.mult_by_i(iref, divide=True), divide=True)), ("div_by_i", Q, P, level)
        # ct x ct, rescale, scale-matched add
        rlk = key(seed + 300)
        c.upload_rlk(rlk, level)
        prod, pref = c.MulRelinNew(A, B), o.mul_relin(a, b, rlk)
        assert eq(prod, pref), ("mul_relin", Q, P, level)
        c.Rescale(prod, PR.SCALE)
        pref = o.rescale(pref, PR.SCALE)
        assert eq(prod, pref), ("rescale", Q, P, level)
        c.Add(prod, B, prod)
        assert eq(prod, o.add_matched(pref, b)), ("add", Q, P, level)
        # plaintext product + constant
        pt = synth.uniform_limbs(seed + 4, Q[:level + 1], N)
        mp = c.MulNew(A, c.upload_pt(pt, PR.SCALE))
        mref = o.mul_pt(a, pt, PR.SCALE)
        assert eq(mp, mref), ("mul_pt", Q, P, level)
        p2 = int(rng.integers(1, 40))
        c.MulByPow2(mp, p2)
        mref = o.mul_by_pow2(mref, p2)
        assert eq(mp, mref), ("mul_by_pow2", Q, P, level, p2)
        const = float(rng.uniform(-3, 3))
        c.MultByConst(mp, const)
        assert eq(mp, o.mul_const(mref, const)), ("mult_by_const", Q, P, level, const)
    finally:
        c.close()


@pytest.mark.parametrize("trial", range(4))
def test_fuzz_conv_path(trial):
    rng = np.random.default_rng(2000 + trial)
    B = int(rng.choice([2, 4, 8, 32]))
    norm = int(rng.choice([n for n in (1, 2, 4) if n <= B]))
    out_scale = float(1 << int(rng.integers(24, 33)))
    Q = (PR.Q_SET6 if trial % 2 == 0 else PR.Q_SET7)[:2]
    c, o = hec.Context(PR.LOGN, Q, PR.P_PACK), Oracle(PR.LOGN, Q, PR.P_PACK)
    try:
        w = synth.conv_workload(Q, PR.P_PACK, PR.LOGN, B, int(rng.integers(1, 1 << 20)))
        idx = o.monomial_pts()
        G = common.GpuConv(c, w, idx, norm)
        bias_pt = c.upload_pt(w["bias"][None, :], out_scale)
        ref = o.conv_then_pack(Ct(*w["ct"][0], PR.SCALE), w["pt_ker"], PR.SCALE, norm, out_scale, idx, w["keys"], w["bias"])[0]
        for flags in (hec.CONV_FUSED, hec.CONV_OPLEVEL):
            assert eq(c.conv_then_pack(G.cts[0], G.ker, norm, out_scale, G.idx, bias_pt, flags), ref), (B, norm, out_scale, flags)
    finally:
        c.close()
