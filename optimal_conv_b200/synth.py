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
