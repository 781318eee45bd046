"""CPU: libhec.so loads, exports every symbol include/hec.h declares, and refuses to
pretend there is a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "hec.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hec_[a-z_0-9]+)\s*\(", src)))


def test_header_and_binding_agree():
    from optimal_conv_b200 import hec
    assert declared_symbols() == sorted(hec.SYMBOLS)


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from optimal_conv_b200 import hec
    L = ctypes.CDLL(hec.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(L, s), s
    L.hec_version.restype = ctypes.c_char_p
    assert b"sm_100a" in L.hec_version()
    hec.lib()


def test_no_cpu_fallback_without_device():
    import torch
    from optimal_conv_b200 import hec, params as PR
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(hec.HecError) as e:
        hec.Context(PR.LOGN, PR.Q_SET6[:2], PR.P_PACK)
    assert e.value.code == hec.HEC_E_CUDA


def test_product_path_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "optimal_conv_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower().replace("no oracle", ""), f


def test_btp_params_struct_layout_matches_the_header(tmp_path):
    """the ctypes mirror of hec_btp_params must have the C compiler's size and field offsets"""
    import ctypes
    import subprocess
    from optimal_conv_b200 import hec
    fields = [f[0] for f in hec.BtpParams._fields_]
    src = tmp_path / "layout.c"
    src.write_text('#include <stddef.h>\n#include <stdio.h>\n#include "hec.h"\nint main(void) { printf("%zu", sizeof(hec_btp_params));\n'
                   + "".join('printf(" %%zu", offsetof(hec_btp_params, %s));\n' % f for f in fields) + "return 0; }\n")
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I" + os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert out[0] == ctypes.sizeof(hec.BtpParams)
    assert out[1:] == [getattr(hec.BtpParams, f).offset for f in fields]


def test_float_quotient_threshold_is_the_reference_step_function():
    """v = uint64(float64(y)/float64(p)) of the single-prime exact basis extension (L:ring/ring_basis_extension.go:670-713)
    is evaluated by the fused mod-down kernel as y >= threshold: check the host-computed threshold against the float
    arithmetic itself on the last 2^16 residues and a seeded sample of the rest, for every special prime and q0."""
    import numpy as np
    import __graft_entry__ as g
    g.build()
    from optimal_conv_b200 import hec, params as PR
    L = hec.lib()
    rng = np.random.default_rng(5)
    for p in PR.P_ALL + [PR.Q_SET6[0], PR.Q_SET6[1], 0x3ffc0001]:
        thr = int(L.hec_float_quotient_threshold(p))
        top = np.arange(p - (1 << 16), p, dtype=np.uint64)
        sample = rng.integers(0, p - (1 << 16), size=1 << 16, dtype=np.uint64)
        for y in (top, sample):
            v = (y.astype(np.float64) / np.float64(p)).astype(np.uint64)
            assert v.max() <= 1
            assert np.array_equal(v == 1, y >= np.uint64(thr) if thr != 2 ** 64 - 1 else np.zeros(len(y), bool)), hex(p)
    assert int(L.hec_float_quotient_threshold(PR.P_ALL[0])) == PR.P_ALL[0] - 129  # the 129-value edge of DESIGN.md 2
