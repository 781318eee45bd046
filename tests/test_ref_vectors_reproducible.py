"""Provenance of tests/golden/ref_*.json: where the reference binary is present (the build container, never the GPU
box) re-interpret a few of its routines and require the committed digests to come out again.  Skipped elsewhere."""
import json
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
BIN = "/root/reference/test_run"
pytestmark = pytest.mark.skipif(not os.path.exists(BIN), reason="reference binary not present (only in the build container)")


def test_interpreted_conv_then_pack_reproduces_the_committed_vector():
    import make_ref_eval_vectors as G
    ref = json.load(open(os.path.join(HERE, "golden", "ref_eval_vectors.json")))
    name, logN, B, norm, seed, out_scale, Q2, P1 = [c for c in G.SMALL_CONV if c[0] == "n8_B2"][0]
    rec = G.conv_case(logN, B, norm, seed, out_scale, Q2, P1)
    assert rec["bias"] == ref["conv"][name]["bias"] and rec["nobias"] == ref["conv"][name]["nobias"]
    assert rec["interpreted_instructions"] == ref["conv"][name]["interpreted_instructions"]


def test_interpreted_linear_transform_reproduces_the_committed_vector():
    import make_ref_eval_vectors as G
    ref = json.load(open(os.path.join(HERE, "golden", "ref_eval_vectors.json")))
    name, *args = G.LT_CASES[-1]
    rec = G.lt_case(*args)
    assert rec["out"] == ref["linear_transform"][name]["out"]


def test_interpreted_generators_reproduce_the_committed_vector():
    import make_ref_vectors as MV
    from optimal_conv_b200 import params as PR
    ref = json.load(open(os.path.join(HERE, "golden", "ref_vectors.json")))
    e = MV.new_emu()
    for q in (PR.Q_SET6[0], PR.Q_SET6[1], PR.P_ALL[0]):
        assert e.call(MV.P + "primitiveRoot", [q, 0])[1] == ref["generator"]["%x" % q]


def test_interpreted_encode_coeffs_and_host_preparation_reproduce_the_committed_vectors():
    import make_ref_eval_vectors as G
    ref = json.load(open(os.path.join(HERE, "golden", "ref_eval_vectors.json")))
    assert G.encode_case("n8_short") == ref["encode_coeffs"]["n8_short"]
    assert G.hostprep_case() == ref["hostprep"]
