"""Pins for the CPU oracle (no GPU).  The reference holds no golden vectors for this
path, so the oracle is pinned by: scalar known answers against Python
big-int arithmetic, table identities, NTT evaluation semantics, schoolbook products,
the automorphism identity, exact big-int rescale / mod-down, and an
encrypt -> conv_then_pack -> decrypt check against a float convolution."""
import random

import numpy as np
import pytest

from optimal_conv_b200 import hostprep as hp
from optimal_conv_b200 import params as PR
from optimal_conv_b200 import synth
from oracle.orc import Ct, Oracle, lib

Q6, P1 = PR.Q_SET6, PR.P_PACK
R = 1 << 64


def brev(x, bits):
    return int(format(x, "0%db" % bits)[::-1], 2)


@pytest.fixture(scope="module")
def orc10():
    return Oracle(10, Q6[:2], P1)


def test_scalar_known_answers():
    L = lib()
    rng = random.Random(1)
    for q in [Q6[0], Q6[1], Q6[2], Q6[5], P1[0]]:
        qinv = pow(q, -1, R)
        b = (1 << 128) // q
        bhi, blo = b >> 64, b & (R - 1)
        for _ in range(200):
            x, y = rng.randrange(q), rng.randrange(q)
            assert L.orc_s_mred(x, y, q, qinv) == x * y * pow(R, -1, q) % q
            assert L.orc_s_mform(x, q, bhi, blo) == x * R % q
            z = rng.randrange(4 * q)
            assert L.orc_s_bred_add(z, q, bhi) == z % q


def test_ring_constants():
    o = Oracle(16, Q6[:2] + Q6[5:6], P1)
    N = 1 << 16
    for ring, limb, q in [(0, 0, Q6[0]), (0, 1, Q6[1]), (0, 2, Q6[5]), (1, 0, P1[0])]:
        assert q % (2 * N) == 1
        assert o.const(ring, limb, 0) == q
        assert o.const(ring, limb, 1) * q % R == 1
        g = o.const(ring, limb, 5)
        # g is the smallest primitive root >= 3 (L:ring/utils.go:69-90)
        m, fac, p = q - 1, [], 2
        while p * p <= m:
            if m % p == 0:
                fac.append(p)
                while m % p == 0:
                    m //= p
            p += 1
        if m > 1:
            fac.append(m)
        isroot = lambda a: all(pow(a, (q - 1) // f, q) != 1 for f in fac)
        assert isroot(g) and not any(isroot(a) for a in range(3, g))
        psi = pow(g, (q - 1) // (2 * N), q)
        assert pow(psi, N, q) == q - 1
        tab, tabi = o.table(ring, limb, 0), o.table(ring, limb, 1)
        rinv = pow(R, -1, q)
        for j in [0, 1, 2, 3, 5, 1000, N - 1]:
            assert int(tab[brev(j, 16)]) * rinv % q == pow(psi, j, q)
            assert int(tabi[brev(j, 16)]) * rinv % q == pow(psi, -j, q)
        assert o.const(ring, limb, 4) == pow(N, -1, q) * R % q


def test_ntt_is_evaluation_at_odd_powers(orc10):
    o, N, q = orc10, 1 << 10, Q6[0]
    a = synth.uniform_mod(5, N, q)
    A = o.ntt(a, 0)
    g = o.const(0, 0, 5)
    psi = pow(g, (q - 1) // (2 * N), q)
    coeffs = [int(x) for x in a]
    for i in [0, 1, 2, 7, 511, 1023]:
        x = pow(psi, 2 * brev(i, 10) + 1, q)
        acc = 0
        for c in reversed(coeffs):
            acc = (acc * x + c) % q
        assert int(A[i]) == acc
    assert np.array_equal(o.intt(A, 0), a)


def test_ntt_product_is_negacyclic_schoolbook():
    o, N = Oracle(6, Q6[:2], P1), 64
    for limb, q in enumerate(Q6[:2]):
        a, b = synth.uniform_mod(1, N, q), synth.uniform_mod(2, N, q)
        A, B = o.ntt(a, limb), o.ntt(b, limb)
        Cn = np.array([int(x) * int(y) % q for x, y in zip(A, B)], dtype=np.uint64)
        c = o.intt(Cn, limb)
        ref = [0] * N
        for i in range(N):
            for j in range(N):
                k, s = (i + j) % N, (-1 if i + j >= N else 1)
                ref[k] = (ref[k] + s * int(a[i]) * int(b[j])) % q
        assert [int(x) for x in c] == ref


def test_full_size_roundtrip_all_moduli():
    allq = PR.Q_SET6 + [PR.Q_SET7[1], PR.Q_SET7[13]]
    o = Oracle(16, allq, PR.P_ALL)
    for ring, mods in ((0, allq), (1, PR.P_ALL)):
        for limb, q in enumerate(mods):
            a = synth.uniform_mod(100 + limb, 1 << 16, q)
            A = o.ntt(a, limb, ring)
            assert A.max() < q
            assert np.array_equal(o.intt(A, limb, ring), a)


def test_automorphism_identity(orc10):
    o, N, q = orc10, 1 << 10, Q6[0]
    a = synth.uniform_mod(9, N, q)
    for g in [5, 25, (1 << 4) + 1, (1 << 10) + 1, 2 * N - 1, o.galois_for_rotation(-3)]:
        b = np.zeros(N, dtype=np.uint64)
        for j in range(N):
            e = g * j % (2 * N)
            b[e % N] = (q - int(a[j])) % q if e >= N else a[j]
        idx = o.permute_index(g)
        assert np.array_equal(o.ntt(a, 0)[idx], o.ntt(b, 0))
    assert o.galois_for_rotation(1) == 5 and o.galois_for_rotation(-1) == pow(5, 2 * N - 1, 2 * N)


def _crt2(r0, r1, q0, q1):
    return (r0 + q0 * ((r1 - r0) * pow(q0, -1, q1) % q1)) % (q0 * q1)


def test_rescale_is_exact_rounded_division(orc10):
    o, N = orc10, 1 << 10
    q0, q1 = Q6[:2]
    x = synth.uniform_limbs(3, [q0, q1], N)
    X = np.stack([o.ntt(x[0], 0), o.ntt(x[1], 1)])
    y = o.intt(o.div_round_last(X)[0], 0)
    for j in range(0, N, 37):
        v = _crt2(int(x[0][j]), int(x[1][j]), q0, q1)
        # round-half-up with half=(q1-1)>>1: floor((v + half)/q1)
        assert int(y[j]) == ((v + ((q1 - 1) >> 1)) // q1) % q0


def test_const_limbs_conv_case(orc10):
    # SetScale constant of conv: 2^-34 scaled by q1
    k, up = orc10.const_limbs(1, 2.0 ** -34)
    assert up == float(Q6[1])
    want = int(np.floor(Q6[1] * 2.0 ** -34 + 0.5))
    assert [int(v) for v in k] == [want % Q6[0], want % Q6[1]]
    k, up = orc10.const_limbs(1, 3.0)
    assert up == 1.0 and [int(v) for v in k] == [3, 3]
    k, up = orc10.const_limbs(1, -2.5)
    w = int(np.floor(2.5 * Q6[1] + 0.5))
    assert [int(v) for v in k] == [Q6[0] - w % Q6[0], Q6[1] - w % Q6[1]]


def test_moddown_is_floor_division_and_float_v_edge(orc10):
    """ModDownSplitNTTPQ == floor((x - [x]_P)/P) on the CRT value, including the
    coefficients where float64(y)/float64(P) rounds up to 1."""
    o, N = orc10, 1 << 10
    q0, p0 = Q6[0], P1[0]
    xq = synth.uniform_mod(11, N, q0)
    xp = synth.uniform_mod(12, N, p0)
    # force the float edge: y in [p0-129, p0-1] gives v = 1, else 0
    xp[0], xp[1], xp[2], xp[3] = p0 - 1, p0 - 129, p0 - 130, p0 - 64
    out = o.intt(o.moddown(o.ntt(xq, 0)[None, :], o.ntt(xp, 0, 1)[None, :])[0], 0)
    pinv = pow(p0, -1, q0)
    for j in list(range(8)) + list(range(8, N, 53)):
        y = int(xp[j])
        v = int(float(y) / float(p0))
        assert v == (1 if y >= p0 - 129 else 0)
        assert int(out[j]) == (int(xq[j]) - (y - v * p0)) * pinv % q0
    # for v = 0 this is floor division of the CRT value by P
    j = 9
    V = _crt2(int(xq[j]), int(xp[j]), q0, p0)
    assert int(out[j]) == (V // p0) % q0


@pytest.mark.parametrize("level", [0, 1])
def test_rotate_gal_decrypts_to_automorphism(level):
    o = Oracle(10, Q6[:2], P1)
    N, q0 = 1 << 10, Q6[0]
    sQ, sP = o.gen_secret(7, 64)
    g = (1 << 5) + 1
    swk = o.gen_rotkey(9, g, sQ, sP)
    vals = np.array([(j * 37) % 1000 - 500 for j in range(N)], dtype=float)
    m = hp.encode_coeffs(vals, 1e6, Q6[:level + 1])
    c0, c1 = o.encrypt(11, m, sQ)
    out = o.rotate_gal(Ct(c0, c1, 1e6), g, swk)
    dec = hp.decode_coeffs(o.decrypt(out, sQ), 1e6, Q6)
    want = np.zeros(N)
    for j in range(N):
        e = g * j % (2 * N)
        want[e % N] = -vals[j] if e >= N else vals[j]
    assert np.abs(dec - want).max() < 1e-3


def test_general_decomposition_alpha2_level1():
    """Baseline shape: alpha=2, level 1 -> one 2-limb digit through the float-assisted
    basis extension."""
    Q = PR.Q_SET7[:2]
    o = Oracle(10, Q, PR.P_PACK_BL)
    N = 1 << 10
    sQ, sP = o.gen_secret(3, 64)
    g = o.galois_for_rotation(3)
    swk = o.gen_rotkey(5, g, sQ, sP)
    vals = np.array([(j * 91) % 777 - 388 for j in range(N)], dtype=float)
    m = hp.encode_coeffs(vals, 2.0 ** 40, Q)
    c0, c1 = o.encrypt(13, m, sQ)
    out = o.rotate_gal(Ct(c0, c1, 2.0 ** 40), g, swk)
    dec = hp.decode_coeffs(o.decrypt(out, sQ), 2.0 ** 40, Q)
    want = np.zeros(N)
    for j in range(N):
        e = g * j % (2 * N)
        want[e % N] = -vals[j] if e >= N else vals[j]
    assert np.abs(dec - want).max() < 1e-6


@pytest.mark.parametrize("B,w,k,norm", [(4, 8, 3, 1), (4, 8, 5, 1), (8, 4, 3, 2)])
def test_conv_then_pack_matches_float_convolution(B, w, k, norm):
    """encrypt -> evalConv_BN hot interval -> decrypt == 'SAME' conv + BN
    (test.go:43-71 with printDebugCfsPlain replaced by an assertion)."""
    N = B * w * w
    logN = N.bit_length() - 1
    o = Oracle(logN, Q6[:2], P1)
    rng = np.random.default_rng(1)
    raw_w = w - k // 2
    rb = B // norm
    raw = rng.normal(size=raw_w * raw_w * rb)
    ker = rng.uniform(-1, 1, size=rb * rb * k * k) / (k * k)
    bn_a, bn_b = rng.uniform(0.5, 1.5, size=rb), rng.uniform(-1, 1, size=rb)
    sQ, sP = o.gen_secret(21, min(64, N // 4))
    swks = {j: o.gen_rotkey(100 + j, (1 << (j + 1)) + 1, sQ, sP) for j in range(logN)}
    scale = PR.SCALE
    m = hp.encode_coeffs(hp.prep_input(raw, raw_w, w, N, norm), scale, Q6[:2])
    c0, c1 = o.encrypt(5, m, sQ)
    kers = hp.prep_ker_coeffs(N, ker, bn_a, w, k, rb, rb, norm)
    pt_ker = np.stack([np.stack([o.ntt(l, i) for i, l in enumerate(hp.encode_coeffs(kc, scale, Q6[:2]))])
                       for kc in kers])
    out_scale = float(1 << 30)
    bias = o.ntt(hp.encode_coeffs(hp.bias_coeffs(N, bn_b, w, norm), out_scale, Q6[:1])[0], 0)
    res, _, _ = o.conv_then_pack(Ct(c0, c1, scale), pt_ker, scale, norm, out_scale, o.monomial_pts(), swks, bias)
    assert res.level == 0 and res.scale == out_scale
    dec = hp.decode_coeffs(o.decrypt(res, sQ), out_scale, Q6)
    if norm == 1:
        got = hp.post_process(dec, raw_w, w)
        want = hp.plain_conv_same(raw, ker, bn_a, bn_b, raw_w, k, B)
    else:
        # only channels i % norm == 0 carry data (conv.go:272,526)
        got = hp.post_process(dec, raw_w, w).reshape(raw_w, raw_w, B)[:, :, ::norm].reshape(-1)
        want = hp.plain_conv_same(raw, ker, bn_a, bn_b, raw_w, k, rb)
    assert np.abs(got - want).max() < 2e-3, np.abs(got - want).max()
