"""CKKS parameter sets used by the reference's conv path.

Extracted from the static data of the reference's prebuilt binary (SURVEY.md
Appendix A; ckks.DefaultBootstrapParams[6] / [7], main.go:52-55) -- the reference
source only names them by index.
"""

# special primes P (61-bit), main.go:416-430,446-454
P_ALL = [0x1fffffffffe00001, 0x1fffffffffc80001, 0x1fffffffffb40001,
         0x1fffffffff500001, 0x1fffffffff420001]

_RELU = [0x3ffc0001, 0x40080001, 0x3fac0001, 0x40720001, 0x3f820001, 0x3f760001,
         0x40980001, 0x3f5a0001, 0x3f540001, 0x40b00001, 0x40c20001]
_STC = [0x1000000000b00001, 0x1000000000ce0001]
_SINE = [0x80000000440001, 0x7fffffffba0001, 0x80000000500001, 0x7fffffffaa0001,
         0x800000005e0001, 0x7fffffff7e0001, 0x7fffffff380001, 0x80000000ca0001]
_CTS = [0x200000000e0001, 0x20000000140001, 0x20000000280001, 0x1fffffffd80001]

# Set 6 ("Ours", main.go:52): Residual | StC | Rotate | ReLU | Sine | CtS
Q_SET6 = [0x80000000080001, 0x1ffffffea0001] + _STC + [0x3ffffe80001] + _RELU + _SINE + _CTS
# Set 7 (baseline, main.go:53-55): Residual(14) | StC | Sine | CtS
Q_SET7 = [0x80000000080001, 0x10000000006e0001] + _RELU + [0xffffffffffc0001] + _STC + _SINE + _CTS

LOGN = 16
ECD_LV = 1                      # main.go:46
SCALE = float(1 << 30)          # params.Scale()

# pack evaluator of "Ours": P = {0x1fffffffffe00001}  (main.go:446-454)
P_PACK = P_ALL[:1]
# pack evaluator of the baseline: two special primes (main.go:416-430)
P_PACK_BL = P_ALL[:2]

# (B, w) table of the `conv` CLI (main.go:578-579)
BATCHES = [4, 16, 64, 256]
WIDTHS = [128, 64, 32, 16]

assert len(Q_SET6) == 28 and len(Q_SET7) == 28
