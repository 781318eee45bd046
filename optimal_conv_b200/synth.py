"""Seeded synthetic CKKS operands: i.i.d. uniform canonical residues
from SplitMix64 with rejection sampling.  Shared by tests and bench; pure numpy."""
import numpy as np

_M64 = (1 << 64) - 1


def splitmix64(seed: int, n: int) -> np.ndarray:
    """n outputs of SplitMix64 started at `seed` (vectorised)."""
    with np.errstate(over="ignore"):
        idx = np.arange(1, n + 1, dtype=np.uint64)
        z = np.uint64(seed & _M64) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def uniform_mod(seed: int, n: int, q: int) -> np.ndarray:
    """n uniform residues in [0,q): mask to bit-length of q, reject >= q."""
    mask = np.uint64((1 << int(q).bit_length()) - 1)
    out = np.empty(n, dtype=np.uint64)
    filled, s = 0, seed
    while filled < n:
        want = n - filled
        draw = splitmix64(s, 2 * want + 16) & mask
        s = (s + 0x632BE59BD9B4E019 * (2 * want + 16)) & _M64
        ok = draw[draw < np.uint64(q)][:want]
        out[filled:filled + len(ok)] = ok
        filled += len(ok)
    return out


def uniform_limbs(seed: int, moduli, n: int) -> np.ndarray:
    """[len(moduli)][n] uniform canonical residues, one stream per limb."""
    return np.stack([uniform_mod(seed * 1000003 + 7919 * i, n, q) for i, q in enumerate(moduli)])


def conv_workload(Q, P, logN, B, seed, n_ct=1, kernels=True):
    """Seeded synthetic operands of one evalConv_BN hot interval:
    n_ct level-1 input ciphertexts, B kernel plaintexts (level 1), a level-0 bias plaintext,
    and rotation keys (uniform words read as NTT+Montgomery form) for the pack levels of B.
    Returns a dict of numpy arrays; no GPU, no oracle."""
    N = 1 << logN
    q2 = list(Q[:2])
    mods = list(Q) + list(P)
    beta_full = (len(Q) + len(P) - 1) // len(P)
    w = {"B": B, "N": N}
    w["ct"] = [(uniform_limbs(seed * 31 + 2 * m, q2, N), uniform_limbs(seed * 31 + 2 * m + 1, q2, N))
               for m in range(n_ct)]
    if kernels:  # kernels=False: only the ciphertexts, the bias and the pack keys of B (wide packings with few real channels)
        w["pt_ker"] = np.stack([uniform_limbs(seed * 977 + 100 + i, q2, N) for i in range(B)])
    w["bias"] = uniform_mod(seed * 131 + 7, N, Q[0])
    keys = {}
    step, log_step = B // 2, 0
    while (1 << log_step) < max(step, 1):
        log_step += 1
    j = logN - log_step
    while step >= 1:
        keys[j - 1] = np.stack([np.stack([uniform_limbs(seed * 7919 + 1000 * j + 10 * d + k, mods, N)
                                          for k in range(2)]) for d in range(beta_full)])
        step //= 2
        j += 1
    w["keys"] = keys  # index i <-> galEl 2^(i+1)+1 (conv.go:255)
    return w


# ---- seeded operands of the split bootstrapping: shared by the golden-vector generator, the
# parity tests and bench.py.  The matrices are synthetic (uniform residues on chosen diagonals): their construction from
# the DFT is the Go host's encoder and stays there; what matters to the evaluation is which diagonals are present.
CTOS_SPECS = [(2, [0, 1, 2, 3, 5], 27), (2, [0, 1, 4, 6], 26), (4, [0, 1, 2, 5], 25), (2, [1, 2, 3], 24)]   # pDFTInv: (N1, diagonals, level)
CTOS_FIELDS = {"prescale": 2.0 ** 47, "postscale": 2.0 ** 47, "sinescale": 2.0 ** 55, "sqrt2pi": 0.3989422804014327, "sc_fac": 4.0,
               "message_ratio": 256.0, "sin_type": 1, "sin_rescal": 2, "params_scale": float(1 << 30)}
BTP_STOC_SPECS = [(2, [0, 1, 2], 15), (2, [0, 1, 3], 14), (2, [1, 2], 13)]   # pDFT for the un-split Bootstrapp


def dft_factor_specs(log_slots=15, depth=4, top_level=27):
    """Diagonal patterns of the bootstrapper's REAL factor matrices: GenCoeffsToSlotsMatrix / GenSlotsToCoeffsMatrix
    merge the log_slots radix-2 butterfly layers of the special FFT into `depth` factors (the reference's parameter sets
    use CtSDepth = 4, StCDepth = 2-3; Appendix A).  A factor that merges m consecutive layers whose smallest butterfly
    stride is 2^s has the 2^(m+1) - 1 diagonals j * 2^s, |j| < 2^m (indices modulo the slot count); N1 is the power of
    two next to sqrt(#diagonals) (MaxN1N2Ratio = 16 permitting).  Returns [(n1, diagonals, level)], top level first."""
    slots = 1 << log_slots
    layers = [log_slots // depth + (1 if i < log_slots % depth else 0) for i in range(depth)]
    specs, s = [], log_slots
    for i, m in enumerate(layers):
        s -= m
        diags = sorted({(j << s) % slots for j in range(-(1 << m) + 1, 1 << m)})
        n1 = 1
        while n1 * n1 < len(diags):
            n1 *= 2
        specs.append((n1 << s if s else n1, diags, top_level - i))
    return specs


def ctos_operands(N, specs=None):
    """seeded operands: the full 28 + 5 modulus chain of set 6, the CoeffsToSlots factor matrices `specs`
    (default: four small ones at the top four levels; scale = the modulus they consume), a degree-63 Chebyshev sine
    polynomial on [-25/4, 25/4], the keys of their rotations, the conjugation key, the relinearisation key"""
    from . import params as PR
    Q, P = PR.Q_SET6, PR.P_ALL
    specs = CTOS_SPECS if specs is None else specs
    beta = (len(Q) + len(P) - 1) // len(P)
    key = lambda s: np.stack([np.stack([uniform_limbs(s + 10 * d + k, list(Q) + list(P), N) for k in range(2)]) for d in range(beta)])  # noqa: E731
    rots = set()
    for n1, diags, _ in specs:
        rots |= {d % n1 for d in diags if d % n1} | {(d // n1) * n1 for d in diags if d // n1}
    keys = {r: key(9000 + 131 * r) for r in sorted(rots)}
    mats = []
    for mi, (n1, diags, ml) in enumerate(specs):
        D = {d: (uniform_limbs(7000 + 100 * mi + d, Q[:ml + 1], N), uniform_limbs(7500 + 100 * mi + d, P, N)) for d in diags}
        mats.append((D, n1, ml, float(Q[ml])))
    rng = np.random.default_rng(77)
    coeffs = [float(x) for x in rng.uniform(-1, 1, 64)]
    b = dict(CTOS_FIELDS, sine_qi=Q[16:24], cheby=(coeffs, -25.0 / 4, 25.0 / 4), mats=mats)
    return keys, key(9900), key(8000), b


def btp_stoc_mats(N):
    from . import params as PR
    Q, P = PR.Q_SET6, PR.P_ALL
    return [({d: (uniform_limbs(7800 + 100 * mi + d, Q[:ml + 1], N), uniform_limbs(7900 + 100 * mi + d, P, N)) for d in diags},
             n1, ml, float(Q[ml])) for mi, (n1, diags, ml) in enumerate(BTP_STOC_SPECS)]
