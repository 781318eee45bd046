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

    def __init__(self, ctx, w, idx_np, norm=1):
        self.ctx, self.w = ctx, w
        B = w["B"]
        self.ker = [ctx.upload_pt(w["pt_ker"][i], PR.SCALE) if i % norm == 0 else None for i in range(B)]
        self.idx = [ctx.upload_pt(idx_np[i:i + 1], 1.0) for i in range(PR.LOGN)]
        self.bias = ctx.upload_pt(w["bias"][None, :], PR.SCALE)
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
