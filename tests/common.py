"""Shared fixtures for the parity tests: the same seeded operands go to the oracle and,
through the C ABI, to the GPU."""
import hashlib

import numpy as np

from optimal_conv_b200 import params as PR
from optimal_conv_b200 import synth

Q2, P1 = PR.Q_SET6[:2], PR.P_PACK


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a, dtype=np.uint64).tobytes())
    return h.hexdigest()


def oracle_conv(o, w, norm, out_scale, idx_np, m=0, bias=True, nthreads=1):
    from oracle.orc import Ct
    res, _, _ = o.conv_then_pack(Ct(*w["ct"][m], PR.SCALE), w["pt_ker"], PR.SCALE, norm, out_scale, idx_np,
                                 w["keys"], w["bias"] if bias else None, nthreads=nthreads)
    return res


class GpuConv:
    """Uploads a synthetic workload once (kernels, monomials, bias, keys)."""

    def __init__(self, ctx, w, idx_np, norm=1, out_scale=PR.SCALE):
        self.ctx, self.w = ctx, w
        B = w["B"]
        self.ker = [ctx.upload_pt(w["pt_ker"][i], PR.SCALE) if i % norm == 0 else None for i in range(B)]
        self.idx = [ctx.upload_pt(idx_np[i:i + 1], 1.0) for i in range(PR.LOGN)]
        self.bias = ctx.upload_pt(w["bias"][None, :], out_scale)  # eval.go:241: NewPlaintext(params, 0, out_scale)
        for j, k in w["keys"].items():
            ctx.upload_swk((1 << (j + 1)) + 1, k, 0)
        self.cts = [ctx.upload_ct(c0, c1, PR.SCALE) for (c0, c1) in w["ct"]]


# seeded configurations whose oracle outputs are pinned in tests/golden/conv_golden.json
GOLDEN_CONFIGS = [
    {"name": "B4_norm1", "B": 4, "norm": 1, "seed": 101, "out_log": 30},
    {"name": "B16_norm1", "B": 16, "norm": 1, "seed": 102, "out_log": 30},
    {"name": "B16_norm2", "B": 16, "norm": 2, "seed": 103, "out_log": 30},
    {"name": "B8_norm1_s25", "B": 8, "norm": 1, "seed": 104, "out_log": 25},
]


def workload(cfg, n_ct=1):
    return synth.conv_workload(Q2, P1, PR.LOGN, cfg["B"], cfg["seed"], n_ct=n_ct)


# ---- EncodeCoeffs operands (conv.go:513-514; shared by the golden generator and the tests)
ENCODE_EDGE = [0.0, -0.0, 1.0, -1.0, 0.5, -0.5, 1.4999999, 2.5, -2.5, 1e-12, -1e-12, 3.7e5, -3.7e5, 1.8e10, -1.8e10,
               2.0 ** 34, -(2.0 ** 34), 2.0 ** 34 + 1, 1e20, -1e20, 0.49999999999999994 / 2 ** 30, 123456.789,
               float(0x3ffc0001) / 2 ** 30, -float(0x3ffc0001) / 2 ** 30, 2.0 ** 33, -(2.0 ** 33), 1.5 * 2 ** 33,
               2.0 ** 34 * (1 - 2.0 ** -53), 3e30, -3e30, 7.25e100, -7.25e100]
ENCODE_CASES = {  # name: (logN, number of values, amplitude, scale, level)
    "n8_scale30": (8, 256, 4.0, float(1 << 30), 2),
    "n8_short": (8, 200, 1e-3, float(1 << 30), 1),
    "n8_scale60": (8, 256, 64.0, float(1 << 60), 2),
    "n10_kernel": (10, 1024, 0.25, float(1 << 30), 1),
}


def encode_values(name):
    """seeded doubles in (-amp, amp) with the edge cases of scaleUpVecExact in front"""
    from optimal_conv_b200 import synth
    logN, n, amp, scale, level = ENCODE_CASES[name]
    u = synth.splitmix64(0xE1C0DE + n + level, n).astype(np.float64) / 2.0 ** 64
    v = (2.0 * u - 1.0) * amp
    k = min(len(ENCODE_EDGE), n)
    v[:k] = ENCODE_EDGE[:k]
    return v
