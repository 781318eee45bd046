"""Pins against the reference's OWN compiled evaluator code, up to main.conv_then_pack itself.

tests/golden/ref_eval_vectors.json was produced by interpreting (never executing) the compiled routines
of the reference's prebuilt binary: the Lattigo fork's ring.NewRing / divRoundByLastModulusNTT /
ModDownSplitNTTPQ, rlwe.SwitchKeysInPlace, ckks.NewEvaluator / Add, and main.conv_then_pack
(tests/golden/make_ref_eval_vectors.py + refmachine.py + x86emu.py).  The CPU tests below run the
oracle on the same seeded operands and require identical SHA-256 digests; the GPU test requires the same
of libhec at N = 2^16."""
import json
import os

import numpy as np
import pytest

import common
from optimal_conv_b200 import params as PR, synth
from oracle.orc import Ct, Oracle

REF = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ref_eval_vectors.json")))
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "conv_golden.json")))


def mods(rec):
    return [int(q, 16) for q in rec["Q"]], [int(p, 16) for p in rec["P"]]


@pytest.mark.parametrize("name", sorted(REF["ring"]))
def test_tables_rescale_moddown_keyswitch_match_reference_code(name):
    """NewRing's tables (psi, psi^-1 bit-reversed Montgomery, N^-1, q^-1 mod 2^64), divRoundByLastModulusNTT
    (every level), ModDownSplitNTTPQ (every level) and SwitchKeysInPlace (every level) for alpha = 1, 2, 5"""
    rec = REF["ring"][name]
    Q, P = mods(rec)
    N = 1 << rec["logN"]
    o = Oracle(rec["logN"], Q, P)
    sel = [(0, i) for i in range(len(Q))] + [(1, i) for i in range(len(P))]
    assert [common.sha(o.table(r, i, 0)) for r, i in sel] == rec["tables"]["psi"]
    assert [common.sha(o.table(r, i, 1)) for r, i in sel] == rec["tables"]["psi_inv"]
    assert [o.const(r, i, 4) for r, i in sel] == rec["tables"]["n_inv"]
    assert [o.const(r, i, 1) for r, i in sel] == rec["tables"]["mred"]
    for level in range(1, len(Q)):
        a = np.stack([synth.uniform_mod(10 + i, N, Q[i]) for i in range(level + 1)])
        assert common.sha(o.div_round_last(a)) == rec["div_round"][str(level)], level
    for level in range(len(Q)):
        aQ = np.stack([synth.uniform_mod(20 + i, N, Q[i]) for i in range(level + 1)])
        aP = np.stack([synth.uniform_mod(30 + i, N, P[i]) for i in range(len(P))])
        assert common.sha(o.moddown(aQ, aP)) == rec["moddown"][str(level)], level
    swk = np.stack([np.stack([synth.uniform_limbs(7000 + 10 * d + k, Q + P, N) for k in range(2)])
                    for d in range(o.beta_full)])
    for level in range(len(Q)):
        c1 = synth.uniform_limbs(41 + level, Q[:level + 1], N)
        d0, d1 = o.keyswitch(c1, swk)
        assert [common.sha(d0), common.sha(d1)] == rec["keyswitch"][str(level)], level


def oracle_conv(rec, bias):
    Q, P = mods(rec)
    o = Oracle(rec["logN"], Q, P)
    w = synth.conv_workload(Q, P, rec["logN"], rec["B"], rec["seed"])
    idx = o.monomial_pts()
    assert common.sha(idx) == rec["monomials"]
    return o.conv_then_pack(Ct(*w["ct"][0], rec["ct_scale"]), w["pt_ker"], rec["pt_scale"], rec["norm"], rec["out_scale"],
                            idx, w["keys"], w["bias"] if bias else None)[0]


@pytest.mark.parametrize("name", sorted(REF["conv"]))
def test_conv_then_pack_matches_reference_main_code(name):
    """main.conv_then_pack (conv.go:522-546, incl. pack_ctxts conv.go:266-300) and the bias Add (eval.go:258),
    interpreted from the reference binary at small N, == the oracle's orc_conv_then_pack, bit for bit"""
    rec = REF["conv"][name]
    if "panic" in rec:
        assert rec["panic"] == "gopanic: LV or scale after conv then pack, inconsistent"
        with pytest.raises(RuntimeError, match="LV or scale after conv then pack, inconsistent"):
            oracle_conv(rec, False)
        return
    for key, bias in (("nobias", False), ("bias", True)):
        r = oracle_conv(rec, bias)
        assert (common.sha(r.c0), common.sha(r.c1), r.scale, r.level) == \
            (rec[key]["c0"], rec[key]["c1"], rec[key]["scale"], rec[key]["level"]), (name, key)


def dg(ct):
    return {"c0": common.sha(ct.c0), "c1": common.sha(ct.c1), "scale": ct.scale, "level": ct.level}


@pytest.mark.parametrize("name", sorted(REF["evalops"]))
def test_baseline_path_evaluator_ops_match_reference_code(name):
    """ckks.evaluator RotateHoisted / RotateNew / MulNew / MulRelinNew / AddNew / SubNew / Add(ct, pt) / Rescale (interpreted),
    the ops of preConv_BL / postConv_BL / evalConv_BN_BL_test (conv.go:133,168-171; eval.go:123,130),
    incl. the alpha = 2 level-1 shape of the baseline convolution == the oracle"""
    rec = REF["evalops"][name]
    Q, P = mods(rec)
    N, level = 1 << rec["logN"], rec["level"]
    o = Oracle(rec["logN"], Q, P)
    lim = lambda seed: synth.uniform_limbs(seed, Q[:level + 1], N)  # noqa: E731
    ct, ct2, pt = Ct(lim(61), lim(62), PR.SCALE), Ct(lim(63), lim(64), PR.SCALE), lim(65)
    assert [o.galois_for_rotation(r) for r in rec["rotations"]] == rec["galois"]
    keys = {g: np.stack([np.stack([synth.uniform_limbs(9000 + 131 * n + 10 * d + k, Q + P, N) for k in range(2)])
                         for d in range(o.beta_full)]) for n, g in enumerate(rec["galois"] + [2 * N - 1])}
    assert dg(o.mult_by_i(ct)) == rec["MultByi"] and dg(o.mult_by_i(ct, divide=True)) == rec["DivByi"]
    assert dg(o.conjugate(ct, keys[2 * N - 1])) == rec["conjugate"]
    for r, g in zip(rec["rotations"], rec["galois"]):
        mine = dg(o.rotate(ct, r, keys[g]))
        assert mine == rec["hoisted"][str(r)], ("hoisted", r)
        if str(r) in rec["rotate_new"]:
            assert mine == rec["rotate_new"][str(r)], ("rotate_new", r)
    assert dg(o.mul_pt(ct, pt, PR.SCALE)) == rec["mul_pt"]
    rlk = np.stack([np.stack([synth.uniform_limbs(8000 + 10 * d + k, Q + P, N) for k in range(2)]) for d in range(o.beta_full)])
    assert dg(o.mul_relin(ct, ct2, rlk)) == rec["mul_relin"]
    assert dg(o.mul_relin(ct, ct, rlk)) == rec["square_relin"]
    assert dg(o.add(ct, ct2)) == rec["add"]
    assert dg(o.sub(ct, ct2)) == rec["sub"]
    scales = {"big_small": (PR.SCALE * 12345.678, PR.SCALE), "small_big": (PR.SCALE, PR.SCALE * 12345.678), "x1.5": (PR.SCALE, PR.SCALE * 1.5)}
    assert len(rec["addsub_scaled"]) == 18
    for key, d in rec["addsub_scaled"].items():   # receiver aliasing does not change the result
        op, tag, _ = key.split(":")
        r = o.add_matched(Ct(ct.c0, ct.c1, scales[tag][0]), Ct(ct2.c0, ct2.c1, scales[tag][1]), sub=(op == "Sub"))
        assert dg(r) == d, key
    assert dg(o.mul_by_pow2(ct, 6)) == rec["mul_by_pow2_6"]
    assert dg(o.add_pt(ct, pt)) == rec["add_pt"]
    assert rec["rescale_err"] is False
    assert dg(o.rescale(Ct(ct.c0, ct.c1, PR.SCALE * float(Q[level])), PR.SCALE)) == rec["rescale"]
    if "rescale2" in rec:  # two divisions by one Rescale call (DivRoundByLastModulusManyNTT, nbRescales = 2)
        r2 = o.rescale(Ct(ct.c0, ct.c1, PR.SCALE * float(Q[level]) * float(Q[level - 1])), PR.SCALE)
        assert r2.level == level - 2 and dg(r2) == rec["rescale2"]


def helper_operands(rec):
    """operands of tests/golden/make_ref_eval_vectors.py::layer_helper_case, regenerated from their seeds"""
    Q, P = mods(rec)
    N, level = 1 << rec["logN"], rec["level"]
    o = Oracle(rec["logN"], Q, P)
    keys = {int(r): np.stack([np.stack([synth.uniform_limbs(9700 + 31 * int(r) + 10 * d + k, Q + P, N) for k in range(2)])
                              for d in range(o.beta_full)]) for r in rec["galois"]}
    lim = lambda seed: synth.uniform_limbs(seed, Q[:level + 1], N)  # noqa: E731
    mask = lambda n, lv=level: synth.uniform_limbs(rec["mask_seed"] + n, Q[:lv + 1], N)  # noqa: E731
    return o, Q, N, level, keys, lim, mask


def test_between_layer_helpers_match_reference_main_code():
    """main.ext_ctxt, main.ext_double_ctxt, main.keep_ctxt (conv.go:347-431) and main.postConv_BL (conv.go:146-178),
    interpreted from the reference binary with a stub encoder (slot encoding is outside the path: the n-th
    plaintext it hands out is a seeded vector), == the oracle's compositions: the MulNew / RotateNew / Add order,
    the plaintext scales the routines choose (q_level, sqrt(q_level), params.Scale()) and the closing Rescale"""
    rec = REF["layer_helpers"]
    o, Q, N, level, keys, lim, mask = helper_operands(rec)
    ct = Ct(lim(61), lim(62), PR.SCALE)
    assert {int(r): g for r, g in rec["galois"].items()} == {int(r): o.galois_for_rotation(int(r)) for r in rec["galois"]}
    # ext_ctxt: masks are handed out in the order the reference walks the map (the interpreter iterates sorted keys)
    e = rec["ext_ctxt"]
    assert e["pt_scales"] == [float(Q[level])] * len(e["rots"])
    r_idx = {r: mask(n) for n, r in enumerate(sorted(e["rots"]))}
    assert dg(o.ext_ctxt(ct, r_idx, float(Q[level]), keys, PR.SCALE)) == e["out"]
    # ext_double_ctxt: both stages use sqrt(q_level) as plaintext scale, one Rescale at the end
    e = rec["ext_double_ctxt"]
    sq = float(np.sqrt(np.float64(Q[level])))
    assert e["pt_scales"] == [sq] * (len(e["m_rots"]) + len(e["r_rots"]))
    m_idx = {r: mask(n) for n, r in enumerate(sorted(e["m_rots"]))}
    r_idx = {r: mask(len(m_idx) + n) for n, r in enumerate(sorted(e["r_rots"]))}
    assert dg(o.ext_double_ctxt(ct, m_idx, r_idx, sq, keys, PR.SCALE)) == e["out"]
    # keep_ctxt
    e = rec["keep_ctxt"]
    assert e["pt_scales"] == [float(Q[level])]
    assert dg(o.keep_ctxt(ct, mask(0), float(Q[level]), PR.SCALE)) == e["out"]
    # postConv_BL: k^2 taps, plaintexts at params.Scale(), no rescale
    e = rec["post_conv_bl"]
    k2 = e["ker_wid"] ** 2
    assert e["pt_scales"] == [PR.SCALE] * k2
    cts = [Ct(lim(70 + 2 * t), lim(71 + 2 * t), PR.SCALE) for t in range(k2)]
    assert dg(o.post_conv_bl(cts, [mask(t) for t in range(k2)], PR.SCALE)) == e["out"]


def test_eval_conv_bn_matches_reference_main_code():
    """main.evalConv_BN (eval.go:224-263) interpreted on a hand-built main.context with a stub encoder: the number, order,
    levels and scales of the plaintexts it encodes (max_bat kernels at ECD_LV / params.Scale(), then the bias at level 0 /
    out_scale), conv_then_pack, the consistency check and the bias Add == the oracle on the same seeded plaintexts"""
    rec = REF["eval_conv_bn"]
    Q, P = mods(rec)
    logN, N = rec["logN"], 1 << rec["logN"]
    o = Oracle(logN, Q, P)
    w = synth.conv_workload(Q, P, logN, rec["max_bat"], rec["seed"])
    idx = o.monomial_pts()
    for name, c in rec["cases"].items():
        assert c["encodes"] == [[1, PR.SCALE]] * rec["max_bat"] + [[0, c["out_scale"]]], name
        pl_ker = np.stack([synth.uniform_limbs(rec["pt_seed"] + i, Q[:2], N) for i in range(rec["max_bat"])])
        bias = synth.uniform_limbs(rec["pt_seed"] + rec["max_bat"], Q[:1], N)[0]
        r = o.conv_then_pack(Ct(*w["ct"][0], PR.SCALE), pl_ker, PR.SCALE, rec["norm"], c["out_scale"], idx, w["keys"], bias)[0]
        assert dg(r) == c["out"], name


def test_eval_conv_bn_relu_new_matches_reference_main_code():
    """main.evalConv_BNRelu_new as the shipped binary has it (kind "Conv", iter = 2), interpreted end to end over the whole
    28 + 5 modulus chain at N = 2^4 on a hand-built main.context and Bootstrapper with a stub encoder (5.6e7 instructions):
    evalConv_BN on the pack evaluator, Scale *= 2^pow, BootstrappConv_CtoS, evalReLU + MulByPow2 per half, keep_ctxt per half,
    BootstrappConv_StoC, Rescale == Oracle.conv_bn_relu, the composition hec_conv_bn_relu is tested against on the GPU"""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_ref_eval_vectors as G
    rec = REF["layer"]
    Q, P = mods(rec)
    logN, N, mb = rec["logN"], 1 << rec["logN"], rec["max_bat"]
    op, o = Oracle(logN, Q, PR.P_PACK), Oracle(logN, Q, P)
    keys, kconj, rlk, b, stoc = G.layer_operands(N)
    w = synth.conv_workload(Q, PR.P_PACK, logN, mb, rec["seed"])
    out_scale = 2.0 ** round(np.log2(float(Q[0])) - (rec["pow"] + 8))               # eval.go:369
    assert rec["encodes"] == [[1, PR.SCALE]] * mb + [[0, out_scale]] + [[4, float(Q[4])]] * 2
    pt = lambda n, lv: synth.uniform_limbs(rec["pt_seed"] + n, Q[:lv + 1], N)       # noqa: E731  (the stub encoder's n-th plaintext)
    pl_ker = np.stack([pt(i, 1) for i in range(mb)])
    r = Oracle.conv_bn_relu(op, o, Ct(w["ct"][0][0][:2], w["ct"][0][1][:2], PR.SCALE), pt_ker=[pl_ker], pt_bias=[pt(mb, 0)[0]],
                            pt_scale=PR.SCALE, norm=[rec["norm"]], out_scale=out_scale, pt_idx=op.monomial_pts(), pack_keys=w["keys"],
                            pow=rec["pow"], alpha=rec["alpha"], iter=2, btp=b, keys=keys, key_conj=kconj, rlk=rlk, stoc_mats=stoc,
                            min_scale=PR.SCALE, keep_mask=[pt(mb + 1, 4), pt(mb + 2, 4)], keep_scale=float(Q[4]))
    assert dg(r) == rec["out"]


def test_pre_conv_bl_matches_reference_main_code():
    """main.preConv_BL (conv.go:120-143), interpreted: the k^2 hoisted rotations i*in_wid + j (negative steps and
    the zero step included) of the baseline convolution == the oracle's rotations, in the reference's order"""
    rec = REF["pre_conv_bl"]
    Q, P = mods(rec)
    N = 1 << rec["logN"]
    o = Oracle(rec["logN"], Q, P)
    ct = Ct(synth.uniform_limbs(61, Q, N), synth.uniform_limbs(62, Q, N), PR.SCALE)
    h = rec["ker_wid"] // 2
    assert rec["rotations"] == [i * rec["in_wid"] + j for i in range(-h, h + 1) for j in range(-h, h + 1)]
    for r, d in zip(rec["rotations"], rec["out"]):
        if r == 0:
            assert dg(ct) == d
            continue
        assert o.galois_for_rotation(r) == rec["galois"][str(r)]
        key = np.stack([np.stack([synth.uniform_limbs(9500 + 17 * (r % 997) + k, Q + P, N) for k in range(2)])])
        assert dg(o.rotate(ct, r, key)) == d, r


@pytest.mark.parametrize("name", sorted(REF["relu"]))
def test_eval_relu_matches_reference_main_code(name):
    """main.evalReLU (conv.go:435-480; three EvaluatePoly + AddConstNew + DropLevel + Mul + Relinearize), interpreted
    from the reference binary, == the oracle's restatement of EvaluatePoly's orchestration, bit for bit"""
    rec = REF["relu"][name]
    Q, P = mods(rec)
    N, level = 1 << rec["logN"], rec["level"]
    o = Oracle(rec["logN"], Q, P)
    rlk = np.stack([np.stack([synth.uniform_limbs(8000 + 10 * d + k, Q + P, N) for k in range(2)]) for d in range(o.beta_full)])
    ct = Ct(synth.uniform_limbs(61, Q[:level + 1], N), synth.uniform_limbs(62, Q[:level + 1], N), PR.SCALE)
    if "panic" in rec:
        with pytest.raises(RuntimeError, match="cannot evaluate"):
            o.eval_relu(ct, rec["alpha"], rlk, PR.SCALE)
        return
    assert dg(o.eval_relu(ct, rec["alpha"], rlk, PR.SCALE)) == rec["out"]


@pytest.mark.parametrize("name", sorted(REF["linear_transform"]))
def test_linear_transform_matches_reference_code(name):
    """ckks.(*evaluator).LinearTransform -> MultiplyByDiagMatrixBSGS (interpreted): hoisted baby rotations kept in
    Q||P, one ModDown per giant step, a second key switch without ModDown, one final ModDown == the oracle"""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    rec = REF["linear_transform"][name]
    Q, P = mods(rec)
    N = 1 << rec["logN"]
    o = Oracle(rec["logN"], Q, P)
    beta = o.beta_full
    n1, diags, level, ml = rec["n1"], rec["diags"], rec["level"], rec["mat_level"]
    rots = sorted({d % n1 for d in diags if d % n1} | {(d // n1) * n1 for d in diags if d // n1})
    keys = {r: np.stack([np.stack([synth.uniform_limbs(9000 + 131 * r + 10 * d + k, Q + P, N) for k in range(2)])
                         for d in range(beta)]) for r in rots}
    ct = Ct(synth.uniform_limbs(61, Q[:level + 1], N), synth.uniform_limbs(62, Q[:level + 1], N), PR.SCALE)
    D = {d: (synth.uniform_limbs(7000 + d, Q[:ml + 1], N), synth.uniform_limbs(7500 + d, P, N)) for d in diags}
    assert dg(o.linear_transform(ct, D, n1, ml, PR.SCALE, keys)) == rec["out"]


def test_coeffs_to_slots_and_slots_to_coeffs_match_reference_code():
    """ckks.CoeffsToSlots / ckks.SlotsToCoeffs (interpreted; the linear halves of BootstrappConv_CtoS / _StoC):
    chains of LinearTransform + Rescale, conjugation, real / imaginary split == the oracle"""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_ref_eval_vectors as G
    rec = REF["dft"]
    N = 1 << rec["logN"]
    o = Oracle(rec["logN"], G.DFT_Q, G.DFT_P)
    keys, kconj, mats, cts = G.dft_operands(N)
    a, b = Ct(*cts[0], PR.SCALE), Ct(*cts[1], PR.SCALE)
    c0, c1 = o.coeffs_to_slots(a, mats, keys, kconj)
    assert [dg(c0), dg(c1)] == rec["coeffs_to_slots"]
    assert dg(o.slots_to_coeffs(a, b, mats, keys)) == rec["slots_to_coeffs"]
    assert dg(o.slots_to_coeffs(b, None, mats, keys)) == rec["slots_to_coeffs_real_only"]


@pytest.mark.parametrize("name", sorted(REF["ctos"]))
def test_bootstrapp_conv_ctos_matches_reference_code(name):
    """ckks.(*Bootstrapper).BootstrappConv_CtoS (interpreted, ~4e7 instructions per case): the first half of the split
    bootstrapping over the whole 28-level chain of set 6 == the oracle, including the returned constant"""
    import struct
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_ref_eval_vectors as G
    rec = REF["ctos"][name]
    N = 1 << rec["logN"]
    Q, P = PR.Q_SET6, PR.P_ALL
    o = Oracle(rec["logN"], Q, P)
    keys, kconj, rlk, b = G.ctos_operands(N)
    lv = rec["level_in"]
    ct = Ct(synth.uniform_limbs(61, Q[:lv + 1], N), synth.uniform_limbs(62, Q[:lv + 1], N), rec["ct_scale"])
    c0, c1, const = o.bootstrapp_conv_ctos(ct, b, keys, kconj, rlk)
    assert [dg(c0), dg(c1)] == rec["out"]
    assert struct.unpack("<Q", struct.pack("<d", const))[0] == rec["const_bits"]


@pytest.mark.parametrize("name", sorted(REF["bootstrapp"]))
def test_bootstrapp_matches_reference_code(name):
    """ckks.(*Bootstrapper).Bootstrapp (interpreted): the un-split bootstrapping the baseline network uses
    (test_BL.go:133) -- the CtoS head, the sine evaluation, SlotsToCoeffs with three factors == the oracle"""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_ref_eval_vectors as G
    rec = REF["bootstrapp"][name]
    N = 1 << rec["logN"]
    Q, P = PR.Q_SET6, PR.P_ALL
    o = Oracle(rec["logN"], Q, P)
    keys, kconj, rlk, b = G.ctos_operands(N)
    lv = rec["level_in"]
    ct = Ct(synth.uniform_limbs(61, Q[:lv + 1], N), synth.uniform_limbs(62, Q[:lv + 1], N), rec["ct_scale"])
    assert dg(o.bootstrapp(ct, b, G.btp_stoc_mats(N), keys, kconj, rlk)) == rec["out"]


def test_sparse_packing_sub_sum_and_repacking_coeffs_to_slots_match_reference_code():
    """Bootstrapper.subSum and ckks.CoeffsToSlots with LogSlots = LogN - 3 (interpreted): the rotate-and-add trace and the
    rotation of the imaginary part into the empty half (one ciphertext returned) == the oracle"""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_ref_eval_vectors as G
    rec = REF["sparse"]
    N = 1 << rec["logN"]
    o = Oracle(rec["logN"], G.DFT_Q, G.DFT_P)
    keys, kconj, mats, (a0, a1), ls = G.sparse_operands(N)
    assert ls == rec["log_slots"]
    ct = Ct(a0, a1, PR.SCALE)
    assert dg(o.sub_sum(ct, ls, keys)) == rec["sub_sum"]
    c0, c1 = o.coeffs_to_slots(ct, mats, keys, kconj, ls)
    assert c1 is None and dg(c0) == rec["coeffs_to_slots"]


def cheby_coeffs(deg):
    rng = np.random.default_rng(1000 + deg)
    co = [float(x) for x in rng.uniform(-1, 1, deg + 1)]
    co[2] = co[5] = 0.0
    return co


@pytest.mark.parametrize("name", sorted(REF["cheby"]))
def test_evaluate_cheby_matches_reference_code(name):
    """ckks.(*evaluator).EvaluateCheby (interpreted; degree 63 = the bootstrapper's sine degree) == the oracle"""
    rec = REF["cheby"][name]
    Q, P = mods(rec)
    N, level = 1 << rec["logN"], rec["level"]
    o = Oracle(rec["logN"], Q, P)
    rlk = np.stack([np.stack([synth.uniform_limbs(8000 + 10 * d + k, Q + P, N) for k in range(2)]) for d in range(o.beta_full)])
    ct = Ct(synth.uniform_limbs(61, Q[:level + 1], N), synth.uniform_limbs(62, Q[:level + 1], N), PR.SCALE)
    assert dg(o.evaluate_poly(ct, cheby_coeffs(rec["degree"]), PR.SCALE, rlk, PR.SCALE, cheby=True)) == rec["out"]


@pytest.mark.parametrize("name", sorted(REF["poly"]))
def test_evaluate_poly_dense_matches_reference_code(name):
    """ckks.(*evaluator).EvaluatePoly (interpreted) with dense real coefficients of degree 5, 12, 31 == the oracle
    (evalReLU only exercises odd polynomials of degree 7 and 13)"""
    rec = REF["poly"][name]
    Q, P = mods(rec)
    N, level = 1 << rec["logN"], rec["level"]
    o = Oracle(rec["logN"], Q, P)
    rlk = np.stack([np.stack([synth.uniform_limbs(8000 + 10 * d + k, Q + P, N) for k in range(2)]) for d in range(o.beta_full)])
    ct = Ct(synth.uniform_limbs(61, Q[:level + 1], N), synth.uniform_limbs(62, Q[:level + 1], N), PR.SCALE)
    assert dg(o.evaluate_poly(ct, cheby_coeffs(rec["degree"]), PR.SCALE, rlk, PR.SCALE)) == rec["out"]


FULL = sorted(REF.get("conv_full", {}))


def test_float_side_host_preparation_matches_reference_code():
    """optimal_conv_b200/hostprep.py (what the semantic encrypt -> conv -> decrypt test builds its operands with) against
    main.reshape_ker / encode_ker_final / prep_Input / post_process of the reference binary"""
    from optimal_conv_b200 import hostprep as hp
    d = REF["hostprep"]
    B, w, k, raw_w, norm = (d["cfg"][x] for x in ("B", "w", "k", "raw_w", "norm"))
    rs = hp.reshape_ker(np.arange(1, B * B * k * k + 1, dtype=float), k * k, B)
    assert np.array_equal(rs, np.array(d["reshape_ker"]))
    for i in range(B):
        assert np.array_equal(hp.encode_ker_final(rs, 0, i, w, B, k), np.array(d["encode_ker_final"][i]))
    N = w * w * B
    raw = np.arange(1, raw_w * raw_w * (B // norm) + 1, dtype=float)
    assert np.array_equal(hp.prep_input(raw, raw_w, w, N, norm), np.array(d["prep_input"]))
    assert np.array_equal(hp.post_process(np.arange(1, N + 1, dtype=float), raw_w, w), np.array(d["post_process"]))


def test_slot_index_maps_match_reference_code():
    """main.gen_keep_vec and main.gen_comprs_fast (rot_util.go:141-174, 498-548; both halves) and main.prt_mat_one_norm
    (main.go:920-939), interpreted, == optimal_conv_b200/hostprep.py.  (The shipped binary predates the *_sparse
    generators; those are restated from the source only.)"""
    from optimal_conv_b200 import hostprep as hp
    d = REF["hostprep"]["index_maps"]
    vs, w, kp = d["vec_size"], d["in_wid"], d["kp_wid"]
    for ul in (0, 1):
        assert hp.gen_keep_vec(vs, w, kp, ul).tolist() == d["keep"][ul]
        m_idx, r_idx = hp.gen_comprs_fast(vs, w, kp, 1, ul)
        for mine, ref in ((m_idx, d["comprs_fast"][ul][0]), (r_idx, d["comprs_fast"][ul][1])):
            assert sorted(mine) == sorted(int(k) for k in ref)
            for rot, mask in mine.items():
                assert mask.tolist() == ref[str(rot)], (ul, rot)
    vec = np.arange(1, 4 * 4 * 8 + 1, dtype=float)
    assert hp.mat_one_norm(vec, 8, 2, 2, 3).tolist() == d["mat_one_norm"]
    # files: what write_txt writes, read_txt reads back bit for bit, in Go's 'e' / shortest-digits form
    import tempfile
    vals = [0.0, 1.0, -1.5, 1e-7, 123456789.125, 3.141592653589793, 2.5e300, 100.0, 0.001]
    with tempfile.TemporaryDirectory() as t:
        hp.write_txt(t + "/x.csv", vals)
        assert open(t + "/x.csv").read().split() == ["0e+00", "1e+00", "-1.5e+00", "1e-07", "1.23456789125e+08",
                                                     "3.141592653589793e+00", "2.5e+300", "1e+02", "1e-03"]
        assert hp.read_txt(t + "/x.csv", len(vals)).tolist() == vals
        with pytest.raises(ValueError, match="input size inconsistent"):
            hp.read_txt(t + "/x.csv", 3)


@pytest.mark.parametrize("name", sorted(common.ENCODE_CASES))
def test_encode_coeffs_to_ntt_matches_reference_code(name):
    """Encoder.EncodeCoeffs (ckks.scaleUpVecExact) + ToNTT (ring.NTTLvl) as the reference's compiled code computes
    them, including the big.Float path above 2^64, Go's float->uint64 conversion, the `q - 0 = q` word a negative
    value rounding to zero leaves behind, and the clearing of the tail"""
    rec = REF["encode_coeffs"][name]
    logN, n, amp, scale, level = common.ENCODE_CASES[name]
    assert (rec["logN"], rec["n"], rec["scale"], rec["level"]) == (logN, n, scale, level)
    o = Oracle(logN, PR.Q_SET6[:level + 1], PR.P_ALL[:1])
    v = common.encode_values(name)
    raw = o.scale_up_vec_exact(v, scale, level)
    assert np.array_equal(raw, o.scale_up_vec_exact_py(v, scale, level))
    k = min(len(common.ENCODE_EDGE), n)
    assert [["%x" % int(x) for x in raw[j, :k]] for j in range(level + 1)] == [r[:k] for r in rec["raw_head"]]
    assert common.sha(raw) == rec["raw"]
    assert common.sha(o.encode_coeffs_ntt(v, scale, level)) == rec["ntt"]
    # the non-canonical word is invisible after the transform
    canon = raw % np.array(o.Q[:level + 1], dtype=np.uint64)[:, None]
    assert common.sha(np.stack([o.ntt(canon[j], j) for j in range(level + 1)])) == rec["ntt"]


@pytest.mark.parametrize("name", FULL)
def test_full_size_golden_fixture_is_the_reference_codes_output(name):
    """N = 2^16: the digests in tests/golden/conv_golden.json (what the GPU parity tests compare libhec with)
    are the ones the reference's own main.conv_then_pack + Add produced on the same operands"""
    rec = REF["conv_full"][name]
    assert rec["logN"] == PR.LOGN
    assert (rec["bias"]["c0"], rec["bias"]["c1"]) == (GOLD["conv"][name]["c0"], GOLD["conv"][name]["c1"])
    assert rec["monomials"] == GOLD["monomials"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", FULL)
def test_gpu_conv_matches_reference_main_code_at_full_size(name):
    """libhec (fused plan and op-level replay) == digests produced by the reference's compiled conv_then_pack"""
    from optimal_conv_b200 import hec
    rec = REF["conv_full"][name]
    Q, P = mods(rec)
    c = hec.Context(PR.LOGN, Q, P)
    try:
        w = synth.conv_workload(Q, P, PR.LOGN, rec["B"], rec["seed"])
        idx = Oracle(PR.LOGN, Q, P).monomial_pts()
        G = common.GpuConv(c, w, idx, rec["norm"], rec["out_scale"])
        for flags in (hec.CONV_FUSED, hec.CONV_OPLEVEL):
            for key, bias in (("nobias", None), ("bias", G.bias)):
                res = c.conv_then_pack(G.cts[0], G.ker, rec["norm"], rec["out_scale"], G.idx, bias, flags)
                g0, g1 = res.download()
                assert (common.sha(g0), common.sha(g1), res.scale, res.level) == \
                    (rec[key]["c0"], rec[key]["c1"], rec[key]["scale"], rec[key]["level"]), (name, key, flags)
    finally:
        c.close()
