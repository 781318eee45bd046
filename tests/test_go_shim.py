"""CPU: the committed cgo shim (go/hec_shim.go) only calls functions include/hec.h declares, with the declared number
of arguments.  There is no Go toolchain in the image, so this is the check that keeps the write-only Go in step with
the C ABI (the ctypes binding is checked the same way in test_abi.py)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_prototypes():
    src = open(os.path.join(ROOT, "include", "hec.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(hec_[a-z_0-9]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        protos[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    return protos


def go_calls(path):
    src = open(path).read()
    src = re.sub(r"//[^\n]*", "", src)
    calls = []
    for m in re.finditer(r"\bC\.(hec_[a-z_0-9]+)\(", src):
        depth, i, nargs, seen = 1, m.end(), 0, False
        while depth:
            ch = src[i]
            if ch in "([{":
                depth += 1
            elif ch in ")]}":
                depth -= 1
            elif ch == "," and depth == 1:
                nargs += 1
            if depth and not ch.isspace():
                seen = True
            i += 1
        calls.append((m.group(1), nargs + 1 if seen else 0))
    return calls


def test_go_shim_calls_match_the_header():
    protos = header_prototypes()
    assert len(protos) > 70
    calls = go_calls(os.path.join(ROOT, "go", "hec_shim.go"))
    assert len(calls) >= 35
    for name, nargs in calls:
        assert name in protos, "go/hec_shim.go calls %s, which include/hec.h does not declare" % name
        assert protos[name] == nargs, "%s: header has %d parameters, the Go call passes %d" % (name, protos[name], nargs)
    # the entry points INTEGRATION.md tells a maintainer to bind are all there
    bound = {n for n, _ in calls}
    for must in ("hec_ctx_create", "hec_swk_upload", "hec_rlk_upload", "hec_pt_upload", "hec_ct_upload", "hec_ct_download",
                 "hec_conv_then_pack", "hec_conv_bl", "hec_encode_coeffs_many", "hec_mul_pt_new", "hec_set_scale",
                 "hec_rotate_gal", "hec_rotate_hoisted", "hec_add_pt", "hec_eval_relu", "hec_mul_relin_new"):
        assert must in bound, must


def test_go_shim_types_are_declared_in_the_header():
    hdr = open(os.path.join(ROOT, "include", "hec.h")).read()
    src = open(os.path.join(ROOT, "go", "hec_shim.go")).read()
    for t in set(re.findall(r"\bC\.(hec_[a-z_]+)\b(?!\()", src)):
        assert re.search(r"\b%s\b" % t, hdr), t
    for k in set(re.findall(r"\bC\.(HEC_[A-Z_]+)\b", src)):
        assert re.search(r"\b%s\b" % k, hdr), k
