"""The deferred-transform dataflow of the fused conv path (hec_kernels.cuh (3)) is the reference's computation: its CPU
model (tests/defer_model.py: pairs (U, e), shifts and sigma_g on coefficients, one transform per output polynomial)
reproduces Oracle.conv_then_pack bit for bit at N = 2^16."""
import numpy as np
import pytest

import common
import defer_model as dm
from optimal_conv_b200 import params as PR, synth
from oracle.orc import Oracle


@pytest.mark.parametrize("B,norm", [(4, 1), (8, 2)])
def test_deferred_dataflow_equals_conv_then_pack(B, norm):
    o = Oracle(PR.LOGN, common.Q2, common.P1)
    idx_np = o.monomial_pts()
    w = synth.conv_workload(common.Q2, common.P1, PR.LOGN, B, seed=700 + B)
    ref = common.oracle_conv(o, w, norm, PR.SCALE, idx_np)
    got = dm.model(o, w, norm, PR.SCALE, idx_np)
    assert np.array_equal(got[0], ref.c0[0]) and np.array_equal(got[1], ref.c1[0])


def test_sigma_on_coefficients_is_the_ntt_domain_permutation():
    """sigma_g acting on coefficients (index map n -> n g mod 2N with a sign) == PermuteNTTWithIndex on the transform"""
    o = Oracle(PR.LOGN, common.Q2, common.P1)
    q0 = common.Q2[0]
    e = synth.uniform_mod(5, 1 << PR.LOGN, q0)
    for j in (9, 13, 16):
        g = (1 << j) + 1
        lhs = o.ntt(np.asarray(dm.sigma_coef(e.astype(object), g, q0), dtype=np.uint64), 0)
        rhs = o.ntt(e, 0)[o.permute_index(g)]
        assert np.array_equal(lhs, rhs), j
