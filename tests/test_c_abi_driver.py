"""The C ABI driven by a plain C program (examples/conv_driver.c): no Python, ctypes or torch between the
caller and libhec.so -- the situation of the reference's Go host binding it through cgo.
CPU: the driver compiles with gcc -Wall -Werror against include/hec.h and links against libhec.so.
GPU: it runs the B4_norm1 golden problem and must reproduce the digests of tests/golden/conv_golden.json,
which are the reference's own conv_then_pack output (tests/test_ref_eval_vectors.py)."""
import json
import os
import struct
import subprocess

import numpy as np
import pytest

import common
from optimal_conv_b200 import hec, params as PR, synth
from oracle.orc import Oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "conv_golden.json")))


def build_driver(tmp_path):
    hec.lib()  # builds libhec.so if needed
    libdir = os.path.dirname(hec.LIB_PATH)
    exe = str(tmp_path / "conv_driver")
    subprocess.check_call(["gcc", "-O2", "-std=c11", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "conv_driver.c"), "-L" + libdir, "-lhec",
                           "-Wl,-rpath," + libdir, "-o", exe])
    return exe


def test_c_driver_compiles_and_links_against_the_abi(tmp_path):
    exe = build_driver(tmp_path)
    out = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
    assert "libhec.so" in out and "not found" not in out.split("libhec.so")[1].splitlines()[0]


def write_problem(path, cfg, idx):
    w = common.workload(cfg)
    Q, P, B = common.Q2, common.P1, cfg["B"]
    bits = lambda d: struct.unpack("<Q", struct.pack("<d", d))[0]  # noqa: E731
    hdr = [0x4845435f434f4e56, PR.LOGN, len(Q), len(P)] + Q + P + [B, cfg["norm"], bits(PR.SCALE), bits(PR.SCALE),
                                                                   bits(float(1 << cfg["out_log"])), 1, len(w["keys"])]
    with open(path, "wb") as f:
        f.write(np.array(hdr, dtype=np.uint64).tobytes())
        c0, c1 = w["ct"][0]
        for a in (c0, c1, w["pt_ker"], idx, w["bias"]):
            f.write(np.ascontiguousarray(a, dtype=np.uint64).tobytes())
        for j, k in sorted(w["keys"].items()):
            f.write(np.array([(1 << (j + 1)) + 1], dtype=np.uint64).tobytes())
            f.write(np.ascontiguousarray(k, dtype=np.uint64).tobytes())


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["fused", "oplevel"])
def test_c_driver_reproduces_the_reference_pinned_golden_conv(tmp_path, mode):
    exe = build_driver(tmp_path)
    cfg = common.GOLDEN_CONFIGS[0]
    idx = Oracle(PR.LOGN, common.Q2, common.P1).monomial_pts()
    prob, res = str(tmp_path / "problem.bin"), str(tmp_path / "result.bin")
    write_problem(prob, cfg, idx)
    r = subprocess.run([exe, prob, res] + (["oplevel"] if mode == "oplevel" else []), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    out = np.fromfile(res, dtype=np.uint64)
    n = 1 << PR.LOGN
    gold = GOLD["conv"][cfg["name"]]
    assert (common.sha(out[:n]), common.sha(out[n:2 * n])) == (gold["c0"], gold["c1"])
    assert struct.unpack("<d", out[2 * n].tobytes())[0] == float(1 << cfg["out_log"]) and int(out[2 * n + 1]) == 0
    assert "kernels launched" in r.stdout
