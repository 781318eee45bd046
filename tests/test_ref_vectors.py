"""Pins against the reference's OWN compiled code: tests/golden/ref_vectors.json was produced by
interpreting, instruction by instruction, the ring routines of the reference's prebuilt binary
(test_lattigo @ eb33b0555aaa; tests/golden/make_ref_vectors.py + x86emu.py).  CPU tests check the
oracle against them; the GPU test checks libhec against the same vectors."""
import hashlib
import json
import os

import numpy as np
import pytest

from optimal_conv_b200 import params as PR, synth
from oracle.orc import Oracle

REF = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ref_vectors.json")))
NTT_CASES = sorted(REF["ntt"])


def sha(vals):
    return hashlib.sha256(np.array(vals, dtype=np.uint64).tobytes()).hexdigest()


def test_generator_of_every_modulus_matches_reference_primitiveRoot():
    """ring.primitiveRoot (interpreted) for all 35 moduli of sets 6/7 + P == the oracle's generator"""
    allq = sorted(set(PR.Q_SET6 + PR.Q_SET7 + PR.P_ALL))
    assert len(REF["generator"]) == len(allq) == 35
    o = Oracle(4, allq, [PR.P_ALL[0]])
    for i, q in enumerate(allq):
        assert o.const(0, i, 5) == REF["generator"]["%x" % q], hex(q)


@pytest.mark.parametrize("case", NTT_CASES)
def test_transforms_match_reference_ntt_code(case):
    """ring.NTTLazy / NTT / InvNTT / InvNTTLazy (interpreted) on a seeded limb == oracle, bit for bit;
    and the lazy output ranges are the ones DESIGN.md states (< 6q forward, < 2q inverse)"""
    q, logN = int(case.split(":")[0], 16), int(case.split(":")[1])
    rec = REF["ntt"][case]
    o = Oracle(logN, [q], [PR.P_ALL[0] if q != PR.P_ALL[0] else PR.P_ALL[1]])
    a = synth.uniform_mod(500 + logN, 1 << logN, q)
    f, i = o.ntt(a, 0), o.intt(a, 0)
    assert sha(f) == rec["ntt"] == rec["ntt_lazy_canonical"]
    assert sha(i) == rec["intt"] == rec["intt_lazy_canonical"]
    assert rec["ntt_lazy_max_over_q"] < 6 and rec["intt_lazy_max_over_q"] < 2


def test_permute_index_matches_reference_code():
    for key, h in REF["permute_index"].items():
        logN, gal = (int(x) for x in key.split(":"))
        o = Oracle(logN, [PR.Q_SET6[0]], PR.P_PACK)
        assert sha(o.permute_index(gal).astype(np.uint64)) == h, key


def test_exact_basis_extension_matches_reference_code():
    """ring.reconstructRNS + ring.multSum (interpreted): the float64 overflow count v and the
    extended residues, for alpha = 1 (mod-down P -> q0, incl. the 129-value float edge), alpha = 2
    (baseline digit -> P) and alpha = 5 (main evaluator digit -> P)"""
    m = REF["modup"][0]
    p0, q0 = int(m["S"][0], 16), int(m["target"], 16)
    cols = [int(c[0]) for c in m["cols"]]
    assert m["v"] == [1 if y >= p0 - 129 else 0 for y in cols]  # the edge predicted in SURVEY.md 7.3-1
    o = Oracle(3, [q0], [p0])
    out = o.intt(o.moddown(np.zeros((1, 8), dtype=np.uint64), o.ntt(np.array(cols, dtype=np.uint64), 0, 1)[None, :])[0], 0)
    x = [(-int(v) * p0) % q0 for v in out]   # mod-down of acc_Q = 0 is -x * P^-1
    assert x == [int(r) for r in m["res_mod_target"]]
    for m in REF["modup"][1:]:
        S, tgt = [int(s, 16) for s in m["S"]], int(m["target"], 16)
        P = PR.P_PACK_BL if len(S) == 2 else PR.P_ALL
        o = Oracle(3, S, P)
        cols = np.array([[int(x) for x in c] for c in m["cols"]], dtype=np.uint64).T.copy()  # [limb][8]
        c1 = np.stack([o.ntt(cols[i], i) for i in range(len(S))])
        _, dP = o.decompose_digit(c1, 0)
        j = P.index(tgt)
        got = [int(v) for v in o.intt(np.array([int(v) % tgt for v in dP[j]], dtype=np.uint64), j, 1)]
        assert got == [int(r) for r in m["res_mod_target"]], m["target"]


@pytest.mark.gpu
def test_gpu_transforms_match_reference_ntt_code():
    from optimal_conv_b200 import hec
    for case in NTT_CASES:
        q, logN = int(case.split(":")[0], 16), int(case.split(":")[1])
        if logN != 16:
            continue
        rec = REF["ntt"][case]
        c = hec.Context(16, [q], [PR.P_ALL[1]])
        try:
            a = synth.uniform_mod(500 + logN, 1 << 16, q)
            assert sha(c.ntt(a, 0)) == rec["ntt"]
            assert sha(c.ntt(a, 0, inverse=True)) == rec["intt"]
        finally:
            c.close()
