"""CPU: the oracle reproduces the committed golden vectors (tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import pytest

import common
from optimal_conv_b200 import params as PR, synth
from oracle.orc import Ct, Oracle

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "conv_golden.json")))


@pytest.fixture(scope="module")
def orc():
    return Oracle(PR.LOGN, common.Q2, common.P1)


@pytest.mark.parametrize("cfg", common.GOLDEN_CONFIGS, ids=lambda c: c["name"])
def test_oracle_conv_matches_golden(orc, cfg):
    idx = orc.monomial_pts()
    assert common.sha(idx) == GOLD["monomials"]
    w = common.workload(cfg)
    r = common.oracle_conv(orc, w, cfg["norm"], float(1 << cfg["out_log"]), idx)
    g = GOLD["conv"][cfg["name"]]
    assert (common.sha(r.c0), common.sha(r.c1), r.scale) == (g["c0"], g["c1"], g["scale"])


def test_oracle_threads_do_not_change_bits(orc):
    cfg = common.GOLDEN_CONFIGS[0]
    w = common.workload(cfg)
    idx = orc.monomial_pts()
    a = common.oracle_conv(orc, w, 1, float(1 << 30), idx, nthreads=1)
    b = common.oracle_conv(orc, w, 1, float(1 << 30), idx, nthreads=4)
    assert np.array_equal(a.c0, b.c0) and np.array_equal(a.c1, b.c1)


def test_oracle_ntt_matches_golden_sample():
    qs = [PR.Q_SET6[0], PR.Q_SET6[1], PR.Q_SET6[4], PR.Q_SET6[9]]
    o = Oracle(PR.LOGN, PR.Q_SET6, PR.P_ALL)
    for q in qs:
        limb = PR.Q_SET6.index(q)
        a = synth.uniform_mod(500 + limb, 1 << PR.LOGN, q)
        g = GOLD["ntt"]["0:%x" % q]
        assert common.sha(o.ntt(a, limb)) == g["fwd"] and common.sha(o.intt(a, limb)) == g["inv"]


def test_conv_scale_panic(orc):
    """conv.go:541-543: an out_scale that SetScale cannot reach at level 0 must be refused."""
    cfg = common.GOLDEN_CONFIGS[0]
    w = common.workload(cfg)
    with pytest.raises(RuntimeError):
        # 2^70: the rescale loop divides 0 times -> result would stay at level 1
        common.oracle_conv(orc, w, 1, float(2 ** 80), orc.monomial_pts())
