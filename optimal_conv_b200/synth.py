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


def conv_workload(Q, P, logN, B, seed, n_ct=1):
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
