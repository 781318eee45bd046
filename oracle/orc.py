"""ctypes wrapper around oracle/liborc.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  Parity: pinned by outputs of the reference's own compiled code (see hec_oracle.h).
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liborc.so")

u64p = C.POINTER(C.c_uint64)


def build(force=False):
    src = os.path.join(_HERE, "hec_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return _LIB


def _p(a):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(u64p)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        L.orc_ctx_new.restype = C.c_void_p
        L.orc_ctx_new.argtypes = [C.c_int, u64p, C.c_int, u64p, C.c_int]
        L.orc_ctx_free.argtypes = [C.c_void_p]
        L.orc_get_const.restype = C.c_uint64
        L.orc_get_const.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.orc_get_table.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, u64p]
        L.orc_beta_full.argtypes = [C.c_void_p]
        for f in ("orc_s_mred",):
            getattr(L, f).restype = C.c_uint64
            getattr(L, f).argtypes = [C.c_uint64] * 4
        L.orc_s_mform.restype = C.c_uint64
        L.orc_s_mform.argtypes = [C.c_uint64] * 4
        L.orc_s_bred_add.restype = C.c_uint64
        L.orc_s_bred_add.argtypes = [C.c_uint64] * 3
        L.orc_ntt.argtypes = [C.c_void_p, C.c_int, C.c_int, u64p, u64p]
        L.orc_intt.argtypes = [C.c_void_p, C.c_int, C.c_int, u64p, u64p]
        L.orc_permute_index.argtypes = [C.c_int, C.c_uint64, C.POINTER(C.c_uint32)]
        L.orc_galois_for_rotation.restype = C.c_uint64
        L.orc_galois_for_rotation.argtypes = [C.c_int, C.c_int]
        L.orc_mul_pt.argtypes = [C.c_void_p, C.c_int] + [u64p] * 5
        L.orc_const_limbs.restype = C.c_double
        L.orc_const_limbs.argtypes = [C.c_void_p, C.c_int, C.c_double, u64p]
        L.orc_mul_const.argtypes = [C.c_void_p, C.c_int, u64p, u64p, u64p]
        L.orc_div_round_last_ntt.argtypes = [C.c_void_p, C.c_int, u64p, u64p]
        L.orc_rescale_count.restype = C.c_int
        L.orc_rescale_count.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.POINTER(C.c_double)]
        L.orc_add.argtypes = [C.c_void_p, C.c_int, u64p, u64p, u64p]
        L.orc_sub.argtypes = [C.c_void_p, C.c_int, u64p, u64p, u64p]
        L.orc_keyswitch.argtypes = [C.c_void_p, C.c_int, u64p, u64p, u64p, u64p]
        L.orc_decompose_digit.argtypes = [C.c_void_p, C.c_int, C.c_int, u64p, u64p, u64p]
        L.orc_keyswitch_nomoddown.argtypes = [C.c_void_p, C.c_int, u64p, u64p, u64p, u64p, u64p, u64p]
        L.orc_poly_mulmont.argtypes = [C.c_void_p, C.c_int, C.c_int, u64p, u64p, u64p, C.c_int]
        L.orc_poly_add.argtypes = [C.c_void_p, C.c_int, C.c_int, u64p, u64p, u64p]
        L.orc_scale_up_vec_exact.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int, C.c_double, C.c_int, u64p]
        L.orc_scale_up_vec_exact.restype = None
        L.orc_moddown.argtypes = [C.c_void_p, C.c_int, u64p, u64p, u64p]
        L.orc_rotate_gal.argtypes = [C.c_void_p, C.c_int, u64p, u64p, C.c_uint64, u64p, u64p, u64p]
        L.orc_conv_then_pack.restype = C.c_int
        L.orc_conv_then_pack.argtypes = [C.c_void_p, u64p, u64p, C.c_double, u64p, C.c_double, C.c_int, C.c_int,
                                         C.c_double, u64p, C.POINTER(u64p), u64p, u64p, u64p,
                                         C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double),
                                         C.POINTER(C.c_double)]
        L.orc_gen_secret.argtypes = [C.c_void_p, C.c_uint64, C.c_int, u64p, u64p]
        L.orc_gen_rotkey.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, u64p, u64p, u64p]
        L.orc_encrypt.argtypes = [C.c_void_p, C.c_uint64, C.c_int, u64p, u64p, u64p, u64p]
        L.orc_decrypt.argtypes = [C.c_void_p, C.c_int, u64p, u64p, u64p, u64p]
        _lib = L
    return _lib


class Ct:
    """Oracle ciphertext: c0,c1 [(level+1)][N] uint64, NTT domain, canonical."""

    def __init__(self, c0, c1, scale):
        self.c0, self.c1, self.scale = np.ascontiguousarray(c0), np.ascontiguousarray(c1), float(scale)

    @property
    def level(self):
        return self.c0.shape[0] - 1


class Oracle:
    def __init__(self, logN, Q, P):
        self.L = lib()
        self.logN, self.N, self.Q, self.P = logN, 1 << logN, list(Q), list(P)
        q = np.array(self.Q, dtype=np.uint64)
        p = np.array(self.P if self.P else [0], dtype=np.uint64)
        self.h = self.L.orc_ctx_new(logN, _p(q), len(self.Q), _p(p), len(self.P))
        self.beta_full = self.L.orc_beta_full(self.h) if self.P else 0

    def __del__(self):
        try:
            self.L.orc_ctx_free(self.h)
        except Exception:
            pass

    # ---- tables ----
    def const(self, ring, limb, which):
        return int(self.L.orc_get_const(self.h, ring, limb, which))

    def table(self, ring, limb, which):
        out = np.empty(self.N, dtype=np.uint64)
        self.L.orc_get_table(self.h, ring, limb, which, _p(out))
        return out

    # ---- ring ----
    def ntt(self, a, limb, ring=0):
        a = np.ascontiguousarray(a, dtype=np.uint64)
        out = np.empty_like(a)
        self.L.orc_ntt(self.h, ring, limb, _p(a), _p(out))
        return out

    def intt(self, a, limb, ring=0):
        a = np.ascontiguousarray(a, dtype=np.uint64)
        out = np.empty_like(a)
        self.L.orc_intt(self.h, ring, limb, _p(a), _p(out))
        return out

    def permute_index(self, galEl):
        idx = np.empty(self.N, dtype=np.uint32)
        self.L.orc_permute_index(self.logN, galEl, idx.ctypes.data_as(C.POINTER(C.c_uint32)))
        return idx

    def galois_for_rotation(self, k):
        return int(self.L.orc_galois_for_rotation(self.logN, k))

    # ---- evaluator ----
    def mul_pt(self, ct, pt, pt_scale=1.0):
        lv = min(ct.level, pt.shape[0] - 1)
        c0, c1, p = (np.ascontiguousarray(x[:lv + 1]) for x in (ct.c0, ct.c1, pt))
        o0, o1 = np.empty_like(c0), np.empty_like(c1)
        self.L.orc_mul_pt(self.h, lv, _p(c0), _p(c1), _p(p), _p(o0), _p(o1))
        return Ct(o0, o1, ct.scale * pt_scale)

    def const_limbs(self, level, constant):
        k = np.zeros(level + 1, dtype=np.uint64)
        up = self.L.orc_const_limbs(self.h, level, constant, _p(k))
        return k, up

    def mul_const(self, ct, constant):
        k, up = self.const_limbs(ct.level, constant)
        o0, o1 = np.empty_like(ct.c0), np.empty_like(ct.c1)
        self.L.orc_mul_const(self.h, ct.level, _p(ct.c0), _p(k), _p(o0))
        self.L.orc_mul_const(self.h, ct.level, _p(ct.c1), _p(k), _p(o1))
        return Ct(o0, o1, ct.scale * up)

    def div_round_last(self, a):
        a = np.ascontiguousarray(a)
        lv = a.shape[0] - 1
        out = np.empty((lv, self.N), dtype=np.uint64)
        self.L.orc_div_round_last_ntt(self.h, lv, _p(a), _p(out))
        return out

    def rescale(self, ct, min_scale):
        """Rescale (L:ckks/evaluator.go:1291-1325); only 0 or 1 division supported here."""
        if ct.level == 0:
            raise ValueError("cannot Rescale: input Ciphertext already at level 0")
        s = C.c_double()
        nb = self.L.orc_rescale_count(self.h, ct.level, ct.scale, min_scale, C.byref(s))
        c0, c1 = ct.c0, ct.c1
        for _ in range(nb):
            c0, c1 = self.div_round_last(c0), self.div_round_last(c1)
        return Ct(c0, c1, s.value)

    def set_scale(self, ct, scale):
        """SetScale (L:ckks/evaluator.go:1194-1209)."""
        t = self.mul_const(ct, scale / ct.scale)
        t = self.rescale(t, scale)
        t.scale = scale
        return t

    def add(self, a, b):
        lv = min(a.level, b.level)
        x0, x1, y0, y1 = (np.ascontiguousarray(v[:lv + 1]) for v in (a.c0, a.c1, b.c0, b.c1))
        o0, o1 = np.empty_like(x0), np.empty_like(x1)
        self.L.orc_add(self.h, lv, _p(x0), _p(y0), _p(o0))
        self.L.orc_add(self.h, lv, _p(x1), _p(y1), _p(o1))
        return Ct(o0, o1, a.scale)

    def sub(self, a, b):
        lv = min(a.level, b.level)
        x0, x1, y0, y1 = (np.ascontiguousarray(v[:lv + 1]) for v in (a.c0, a.c1, b.c0, b.c1))
        o0, o1 = np.empty_like(x0), np.empty_like(x1)
        self.L.orc_sub(self.h, lv, _p(x0), _p(y0), _p(o0))
        self.L.orc_sub(self.h, lv, _p(x1), _p(y1), _p(o1))
        return Ct(o0, o1, a.scale)

    def mul_relin(self, a, b, rlk):
        """MulRelinNew(ct, ct) (mulRelin ct branch, L:ckks/evaluator.go:1398-1432): c0 = a0 b0,
        c1 = a0 b1 + a1 b0, c2 = a1 b1, then (c0, c1) += SwitchKeys(c2, rlk).  Composition of the
        pinned routines above (MForm + MulCoeffsMontgomery = canonical product)."""
        lv = min(a.level, b.level)
        a = Ct(a.c0[:lv + 1], a.c1[:lv + 1], a.scale)
        x = self.mul_pt(a, b.c0[:lv + 1])            # a0 b0, a1 b0
        y = self.mul_pt(a, b.c1[:lv + 1])            # a0 b1, a1 b1
        c1 = np.empty_like(x.c0)
        self.L.orc_add(self.h, lv, _p(y.c0), _p(x.c1), _p(c1))
        d0, d1 = self.keyswitch(y.c1, rlk)
        o0, o1 = np.empty_like(c1), np.empty_like(c1)
        self.L.orc_add(self.h, lv, _p(x.c0), _p(d0), _p(o0))
        self.L.orc_add(self.h, lv, _p(c1), _p(d1), _p(o1))
        return Ct(o0, o1, a.scale * b.scale)

    # ---- polynomial evaluation (evalReLU, conv.go:435-480; L:ckks/polynomial_evaluation.go) ----
    @staticmethod
    def scale_up_exact(value, n, q):
        """scaleUpExact (L:ckks/utils.go:31-57): big.Float at 53 bits == IEEE double; Int truncates"""
        neg = value < 0
        x = -n * value if neg else n * value
        r = int(x + 0.5) % q
        return q - r if neg else r

    # ---- Encoder.EncodeCoeffs + ToNTT (conv.go:513-514, eval.go:242-243; L:ckks/encoder.go, L:ckks/utils.go) ----
    @staticmethod
    def _f2u(x):
        """Go's float64 -> uint64 conversion as compiled for amd64 (CVTTSD2SI, split at 2^63)"""
        def cvt(y):
            return (1 << 63) if (y != y or y >= 2.0 ** 63 or y < -2.0 ** 63) else int(y) & ((1 << 64) - 1)
        return cvt(x) if x < 2.0 ** 63 else cvt(x - 2.0 ** 63) | (1 << 63)

    def scale_up_vec_exact(self, values, n, level):
        """scaleUpVecExact in C (orc_scale_up_vec_exact); scale_up_vec_exact_py is the same in big-integer Python"""
        v = np.ascontiguousarray(values, dtype=np.float64)
        out = np.empty((level + 1, self.N), dtype=np.uint64)
        self.L.orc_scale_up_vec_exact(self.h, v.ctypes.data_as(C.POINTER(C.c_double)), v.shape[0], n, level, _p(out))
        return out

    def scale_up_vec_exact_py(self, values, n, level):
        """scaleUpVecExact (L:ckks/utils.go:59-123) as the pinned binary computes it: |n*v| > 2^64 goes through
        big.Float (53 bits == IEEE double) and big.Int; otherwise uint64(n*v + 0.5) % q; negative values give
        q - r, which is q itself (not 0) when r == 0; coefficients past len(values) are cleared."""
        out = np.zeros((level + 1, self.N), dtype=np.uint64)
        for i, v in enumerate(np.asarray(values, dtype=np.float64)):
            v, neg = float(v), float(v) < 0
            if n * abs(v) > 2.0 ** 64:
                x = int((-n * v if neg else n * v) + 0.5)
            else:
                x = self._f2u((-n * v if neg else n * v) + 0.5)
            for j in range(level + 1):
                r = x % self.Q[j]
                out[j, i] = self.Q[j] - r if neg else r
        return out

    def encode_coeffs_ntt(self, values, scale, level):
        """EncodeCoeffs then ToNTT: the plaintext limbs prep_Ker / evalConv_BN hand to the evaluator"""
        raw = self.scale_up_vec_exact(values, scale, level)
        return np.stack([self.ntt(raw[j], j) for j in range(level + 1)])

    # ---- hoisted linear transform (CoeffsToSlots / SlotsToCoeffs of the bootstrapper; L:ckks/linear_transform.go) ----
    def keyswitch_nomoddown(self, c1, swk):
        c1 = np.ascontiguousarray(c1)
        lv, nP = c1.shape[0] - 1, len(self.P)
        a0Q, a1Q = np.empty_like(c1), np.empty_like(c1)
        a0P, a1P = np.empty((nP, self.N), dtype=np.uint64), np.empty((nP, self.N), dtype=np.uint64)
        self.L.orc_keyswitch_nomoddown(self.h, lv, _p(c1), _p(np.ascontiguousarray(swk)), _p(a0Q), _p(a0P), _p(a1Q), _p(a1P))
        return a0Q, a0P, a1Q, a1P

    def _mulmont(self, ring, a, b, out=None):
        a, b = np.ascontiguousarray(a), np.ascontiguousarray(b[:a.shape[0]])
        acc = out is not None
        out = out if acc else np.empty_like(a)
        self.L.orc_poly_mulmont(self.h, ring, a.shape[0], _p(a), _p(b), _p(out), int(acc))
        return out

    def _padd(self, ring, a, b):
        a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
        out = np.empty_like(a)
        self.L.orc_poly_add(self.h, ring, a.shape[0], _p(a), _p(b), _p(out))
        return out

    def linear_transform(self, ct, diags, n1, mat_level, mat_scale, keys):
        """LinearTransform(ct, *PtDiagMatrix) -> MultiplyByDiagMatrixBSGS (L:ckks/linear_transform.go): baby-step
        giant-step product of the slot vector with a matrix given by its non-zero diagonals.
          diags: {k: (dQ [(level+1)][N], dP [nP][N])}, the plaintext diagonals over Q and P in Montgomery form
                 (PtDiagMatrix.Vec), k in [0, slots); n1: PtDiagMatrix.N1 (power of two);
          keys:  {rotation r: switching key for galEl 5^r} for the baby steps k mod n1 and giant steps n1*(k // n1).
        Structure (every ModDown is a rounding step, so its place decides bits): the baby rotations stay in Q||P
        (KeyswitchHoistedNoModDown + P*c0), each giant step's inner sum is ModDown'ed, key-switched again without
        ModDown and accumulated in Q||P, one final ModDown; the un-rotated terms are added in Q."""
        lv = min(ct.level, mat_level)
        L, nP = lv + 1, len(self.P)
        c0, c1 = np.ascontiguousarray(ct.c0[:L]), np.ascontiguousarray(ct.c1[:L])
        index = {}
        for k in sorted(diags):
            index.setdefault(k // n1, []).append(k & (n1 - 1))
        # rotateHoistedNoModDown: R_i = perm_i(KS_noModDown(c1, key_i) + (P c0, 0)) over Q||P
        Pmod = 1
        for p in self.P:
            Pmod *= p
        pc0 = np.empty_like(c0)
        self.L.orc_mul_const(self.h, lv, _p(c0), _p(np.array([Pmod % self.Q[i] for i in range(L)], dtype=np.uint64)), _p(pc0))
        rot = {}
        for i in sorted({i for v in index.values() for i in v if i}):
            a0Q, a0P, a1Q, a1P = self.keyswitch_nomoddown(c1, keys[i])
            a0Q = self._padd(0, a0Q, pc0)
            idx = self.permute_index(self.galois_for_rotation(i))
            rot[i] = tuple(np.ascontiguousarray(x[:, idx]) for x in (a0Q, a0P, a1Q, a1P))
        res0, res1 = np.zeros_like(c0), np.zeros_like(c1)
        outer = None
        for j in sorted(index):
            acc = None
            for i in index[j]:
                if i == 0:
                    continue
                dQ, dP = diags[n1 * j + i]
                terms = (self._mulmont(0, rot[i][0], dQ), self._mulmont(1, rot[i][1], dP),
                         self._mulmont(0, rot[i][2], dQ), self._mulmont(1, rot[i][3], dP))
                acc = terms if acc is None else tuple(self._padd(r, a, t) for r, a, t in zip((0, 1, 0, 1), acc, terms))
            if j == 0:
                if acc is not None:
                    outer = acc if outer is None else tuple(self._padd(r, a, t) for r, a, t in zip((0, 1, 0, 1), outer, acc))
                continue
            if acc is not None:
                t0, t1 = self.moddown(acc[0], acc[1]), self.moddown(acc[2], acc[3])
            else:
                t0, t1 = np.zeros_like(c0), np.zeros_like(c1)
            if 0 in index[j]:
                dQ = diags[n1 * j][0]
                t0, t1 = self._mulmont(0, c0, dQ, t0), self._mulmont(0, c1, dQ, t1)
            idx = self.permute_index(self.galois_for_rotation(n1 * j))
            d = self.keyswitch_nomoddown(t1, keys[n1 * j])
            res0 = self._padd(0, res0, np.ascontiguousarray(t0[:, idx]))
            d = tuple(np.ascontiguousarray(x[:, idx]) for x in d)
            outer = d if outer is None else tuple(self._padd(r, a, t) for r, a, t in zip((0, 1, 0, 1), outer, d))
        if outer is not None:
            res0 = self._padd(0, res0, self.moddown(outer[0], outer[1]))
            res1 = self._padd(0, res1, self.moddown(outer[2], outer[3]))
        if 0 in index.get(0, []):
            dQ = diags[0][0]
            res0, res1 = self._mulmont(0, c0, dQ, res0), self._mulmont(0, c1, dQ, res1)
        return Ct(res0, res1, ct.scale * mat_scale)

    def dft(self, ct, mats, keys):
        """dft (L:ckks/bootstrap.go): the chain of LinearTransform + Rescale(back to the scale before) over the
        factor matrices.  mats: [(diags, n1, level, scale)]"""
        for diags, n1, ml, ms in mats:
            s = ct.scale
            ct = self.rescale(self.linear_transform(ct, diags, n1, ml, ms, keys), s)
        return ct

    def coeffs_to_slots(self, ct, mats, keys, key_conj, log_slots=None):
        """CoeffsToSlots (L:ckks/bootstrap.go): z = dft(ct); real part z + conj(z), imaginary part (z - conj(z)) / i.
        Sparse packing (log_slots < logN - 1): the imaginary part is rotated by `slots` into the empty half of the
        real part and one ciphertext is returned (ct1 = None)."""
        z = self.dft(ct, mats, keys)
        zc = self.conjugate(z, key_conj)
        c0, c1 = self.add_matched(z, zc), self.mult_by_i(self.add_matched(z, zc, sub=True), divide=True)
        if log_slots is not None and log_slots < self.logN - 1:
            return self.add_matched(c0, self.rotate(c1, 1 << log_slots, keys[1 << log_slots])), None
        return c0, c1

    def sub_sum(self, ct, log_slots, keys):
        """Bootstrapper.subSum: ct += Rotate(ct, 2^i) for i = log_slots .. logN - 2 (nothing at full packing)"""
        for i in range(log_slots, self.logN - 1):
            ct = self.add_matched(ct, self.rotate(ct, 1 << i, keys[1 << i]))
        return ct

    def slots_to_coeffs(self, ct0, ct1, mats, keys):
        """SlotsToCoeffs (L:ckks/bootstrap.go): dft(ct0 + i ct1)"""
        if ct1 is not None:
            ct0 = self.add_matched(ct0, self.mult_by_i(ct1))
        return self.dft(ct0, mats, keys)

    # ---- split bootstrapping, first half (BootstrappConv_CtoS of the fork's ckks/bootstrap.go) ----
    def mod_up(self, ct):
        """Bootstrapper.modUp: the level-0 ciphertext's coefficients, centred around q0, re-expressed modulo every
        q_i of the chain (then back to the NTT domain)"""
        q0 = self.Q[0]
        out = []
        for poly in (ct.c0, ct.c1):
            co = self.intt(poly[0], 0)
            neg = co > np.uint64(q0 >> 1)
            limbs = []
            for i, qi in enumerate(self.Q):
                pos = co % np.uint64(qi)
                t = (np.uint64(q0) - co) % np.uint64(qi)          # only meaningful where neg
                v = np.where(neg, (np.uint64(qi) - t) % np.uint64(qi), pos).astype(np.uint64)
                limbs.append(self.ntt(v, i))
            out.append(np.stack(limbs))
        return Ct(out[0], out[1], ct.scale)

    def btp_evaluate_cheby(self, ct, b, rlk):
        """Bootstrapper.evaluateCheby: change of variable, Chebyshev evaluation of the scaled cosine, SinRescal double-angle
        steps (2 x^2 - sqrt2pi^(2^k)); the evaluator's rescale threshold is the sine scale throughout"""
        target = b["sinescale"]
        for i in range(b["sin_rescal"]):
            target = float(np.sqrt(target * float(b["sine_qi"][i])))
        co, lo, hi = b["cheby"]
        if b["sin_type"] in (1, 2):
            ct = self.add_const(ct, -0.5 / (b["sc_fac"] * (hi - lo)))
        ct = self.evaluate_poly(ct, co, target, rlk, b["sinescale"], cheby=True)
        s = b["sqrt2pi"]
        for _ in range(b["sin_rescal"]):
            s *= s
            ct = self.mul_relin(ct, ct, rlk)
            ct = self.add_matched(ct, ct)
            ct = self.rescale(self.add_const(ct, -s), b["sinescale"])
        return ct

    def _btp_until_sine(self, ct, b, keys, key_conj, rlk):
        """common head of Bootstrapp / BootstrappConv_CtoS: to the bootstrapping scale at level 0, modUp, ScaleUp,
        CoeffsToSlots, evaluateSine on both halves"""
        while ct.level > 1:
            ct = self.drop_level(ct, 1)
        if ct.level == 1:
            ct = self.set_scale(ct, b["prescale"])
            ct = self.drop_level(ct, ct.level)
        else:
            if ct.scale > b["prescale"]:
                raise RuntimeError("ciphetext scale > q/||m||)")
            r = float(np.round(b["prescale"] / ct.scale))
            ct = self.mul_const(ct, r)
            ct.scale = ct.scale * r
        ct = self.mod_up(ct)
        r = float(np.round(b["postscale"] / ct.scale))
        ct = self.mul_const(ct, r)
        ct.scale = ct.scale * r
        ls = b.get("log_slots", self.logN - 1)
        ct = self.sub_sum(ct, ls, keys)
        ct0, ct1 = self.coeffs_to_slots(ct, b["mats"], keys, key_conj, ls)
        outs = []
        for c in (ct0, ct1):
            if c is None:          # sparse packing: one ciphertext carries both parts
                outs.append(None)
                continue
            c = Ct(c.c0, c.c1, c.scale * b["message_ratio"])
            c = self.btp_evaluate_cheby(c, b, rlk)
            c.scale = c.scale / (b["postscale"] * b["message_ratio"] / b["params_scale"])
            outs.append(c)
        return outs

    def bootstrapp(self, ct, b, stoc_mats, keys, key_conj, rlk):
        """Bootstrapper.Bootstrapp (test_BL.go:133; 0x505d00 in the reference binary): the head above, then
        SlotsToCoeffs with the pDFT factors; the fork returns that result as it is"""
        c0, c1 = self._btp_until_sine(ct, b, keys, key_conj, rlk)
        return self.slots_to_coeffs(c0, c1, stoc_mats, keys)

    def bootstrapp_conv_ctos(self, ct, b, keys, key_conj, rlk):
        """BootstrappConv_CtoS: scale to the bootstrapping scale at level 0, modUp, CoeffsToSlots, sine evaluation on the
        real and imaginary halves, and the fork's final constant multiplication + Rescale.  b: the Bootstrapper's fields
        (prescale, postscale, sinescale, sqrt2pi, sc_fac, message_ratio, sin_type, sin_rescal, sine_qi, cheby = (coeffs,
        a, b), mats = pDFTInv factors, params_scale).  Returns (ct0, ct1, constant)."""
        outs = self._btp_until_sine(ct, b, keys, key_conj, rlk)
        q0 = float(self.Q[0])
        const = q0 / float(2.0 ** np.round(np.log2(q0))) * b["params_scale"] / b["postscale"]
        # (under sparse packing the reference binary dereferences the nil second half here and dies; the one
        # ciphertext there is gets the same treatment as at full packing)
        outs = [None if c is None else self.rescale(self.mul_const(c, const), b["params_scale"]) for c in outs]
        return outs[0], outs[1], const

    def mult_by_i(self, ct, divide=False):
        """MultByi / DivByi (L:ckks/evaluator.go): product with X^(N/2) in the NTT domain = first half of the slots
        times psi^(N/2) (NttPsi[i][1]), second half times its negative; DivByi swaps the two"""
        lv, h = ct.level, self.N // 2
        R = 1 << 64
        k = np.array([int(self.table(0, i, 0)[1]) * pow(R, -1, self.Q[i]) % self.Q[i] for i in range(lv + 1)], dtype=np.uint64)
        kn = np.array([self.Q[i] - int(k[i]) for i in range(lv + 1)], dtype=np.uint64)
        out = []
        for src in (ct.c0, ct.c1):
            a, b = np.empty_like(src), np.empty_like(src)
            self.L.orc_mul_const(self.h, lv, _p(np.ascontiguousarray(src)), _p(k), _p(a))
            self.L.orc_mul_const(self.h, lv, _p(np.ascontiguousarray(src)), _p(kn), _p(b))
            first, second = (b, a) if divide else (a, b)
            out.append(np.concatenate([first[:, :h], second[:, h:]], axis=1))
        return Ct(out[0], out[1], ct.scale)

    def conjugate(self, ct, swk):
        """Conjugate (L:ckks/evaluator.go): automorphism X -> X^(2N-1) (GaloisElementForRowRotation)"""
        return self.rotate_gal(ct, 2 * self.N - 1, swk)

    def mul_by_pow2(self, ct, pow2):
        """MulByPow2(ct, pow2, ct) (ring.MulByPow2Lvl): coefficients times 2^pow2 mod q_i, scale unchanged"""
        k = np.array([pow(2, pow2, self.Q[i]) for i in range(ct.level + 1)], dtype=np.uint64)
        o0, o1 = np.empty_like(ct.c0), np.empty_like(ct.c1)
        self.L.orc_mul_const(self.h, ct.level, _p(np.ascontiguousarray(ct.c0)), _p(k), _p(o0))
        self.L.orc_mul_const(self.h, ct.level, _p(np.ascontiguousarray(ct.c1)), _p(k), _p(o1))
        return Ct(o0, o1, ct.scale)

    def drop_level(self, ct, levels):
        return Ct(ct.c0[:ct.level + 1 - levels], ct.c1[:ct.level + 1 - levels], ct.scale)

    def zero_ct(self, level, scale):
        return Ct(np.zeros((level + 1, self.N), dtype=np.uint64), np.zeros((level + 1, self.N), dtype=np.uint64), scale)

    def add_const(self, ct, c):
        """AddConst(ct, c real, ct) (L:ckks/evaluator.go AddConst): c0 += scaleUpExact(c, ct.Scale, q_i) in every slot"""
        c0 = ct.c0.copy()
        if c != 0:
            for i in range(ct.level + 1):
                q = self.Q[i]
                k = self.scale_up_exact(c, ct.scale, q) % q
                c0[i] = ((c0[i].astype(object) + k) % q).astype(np.uint64)
        return Ct(c0, ct.c1, ct.scale)

    def mul_int_and_add(self, ct, c_int, out):
        """MultByGaussianIntegerAndAdd(ct, c, 0, out): out += ct * c (L:ckks/evaluator.go), c an int64"""
        lv = min(ct.level, out.level)
        k = np.array([c_int % self.Q[i] for i in range(lv + 1)], dtype=np.uint64)
        o = []
        for src, dst in ((ct.c0, out.c0), (ct.c1, out.c1)):
            t = np.empty((lv + 1, self.N), dtype=np.uint64)
            self.L.orc_mul_const(self.h, lv, _p(np.ascontiguousarray(src[:lv + 1])), _p(k), _p(t))
            r = np.empty_like(t)
            self.L.orc_add(self.h, lv, _p(np.ascontiguousarray(dst[:lv + 1])), _p(t), _p(r))
            o.append(r)
        return Ct(o[0], o[1], out.scale)

    def add_matched(self, a, b, sub=False):
        """Add / Sub(a, b, out) with evaluateInPlace's scale matching (L:ckks/evaluator.go:365-473): the operand
        with the smaller scale is multiplied by floor(ratio) when that is > 1 (MultByConst with an integer
        constant); level = min, scale = max -- whichever of a, b, or a third ciphertext receives the result."""
        lv = min(a.level, b.level)
        a, b = self.drop_level(a, a.level - lv), self.drop_level(b, b.level - lv)
        if a.scale > b.scale and np.floor(a.scale / b.scale) > 1:
            b = self.mul_const(b, float(np.floor(a.scale / b.scale)))
        elif b.scale > a.scale and np.floor(b.scale / a.scale) > 1:
            a = self.mul_const(a, float(np.floor(b.scale / a.scale)))
        r = self.sub(a, b) if sub else self.add(a, b)
        r.scale = max(a.scale, b.scale)
        return r

    def evaluate_poly(self, ct, coeffs, target_scale, rlk, eval_scale, cheby=False):
        """EvaluatePoly (L:ckks/polynomial_evaluation.go: computePowerBasis, recurse, splitCoeffs,
        evaluatePolyFromPowerBasis) and, with cheby=True, EvaluateCheby (computePowerBasisCheby: T_n = 2 T_a T_b -
        T_|a-b|; splitCoeffsCheby).  coeffs: real coefficients, index = degree.  eval_scale = params.Scale()."""
        Q = self.Q
        deg = len(coeffs) - 1
        log_degree = deg.bit_length()
        log_split = log_degree >> 1
        if ct.level < log_degree:
            raise RuntimeError("%d levels < %d log(d) -> cannot evaluate" % (ct.level, log_degree))
        C = {1: Ct(ct.c0.copy(), ct.c1.copy(), ct.scale)}

        def power(n):
            if n not in C:
                a, b = (n + 1) // 2, n >> 1
                power(a)
                power(b)
                if cheby and a != b:
                    power(a - b)
                C[n] = self.rescale(self.mul_relin(C[a], C[b], rlk), eval_scale)
                if cheby:
                    C[n] = self.add_matched(C[n], C[n])                      # 2 T_a T_b
                    C[n] = self.add_const(C[n], -1.0) if a == b else self.add_matched(C[n], C[a - b], sub=True)

        for i in range(2, 1 << log_split):
            power(i)
        for i in range(log_split, log_degree):
            power(1 << i)

        class P:  # Poly{maxDeg, coeffs, lead}
            def __init__(self, c, max_deg, lead):
                self.c, self.max_deg, self.lead = c, max_deg, lead

            @property
            def degree(self):
                return len(self.c) - 1

        def split(p, sp):
            r = P(list(p.c[:sp]), sp - 1 if p.max_deg == p.degree else p.max_deg - (p.degree - sp + 1), False)
            q = P(list(p.c[sp:]), p.max_deg, p.lead)
            if cheby:   # p = q T_sp + r with T_i T_sp = (T_(sp+i) + T_(sp-i)) / 2
                for i in range(sp + 1, p.degree + 1):
                    q.c[i - sp] = 2 * p.c[i]
                    r.c[sp - (i - sp)] -= p.c[i]
            return q, r

        def from_basis(ts, p):
            if p.degree == 0:
                res = self.zero_ct(C[1].level, ts)
                return self.add_const(res, p.c[0]) if abs(p.c[0]) > 1e-14 else res
            lvl = C[p.degree].level
            qi = float(Q[lvl])
            res = self.zero_ct(lvl, ts * qi)
            if abs(p.c[0]) > 1e-14:
                res = self.add_const(res, p.c[0])
            for key in range(p.degree, 0, -1):
                if abs(p.c[key]) > 1e-14:
                    const_scale = ts * qi / C[key].scale
                    x = p.c[key] * const_scale   # int64(float): CVTTSD2SQ yields -2^63 when out of range
                    res = self.mul_int_and_add(C[key], int(x) if abs(x) < 2.0 ** 63 else -(1 << 63), res)
            return self.rescale(res, eval_scale)

        def recurse(ts, ls, ld, p):
            if p.degree < (1 << ls):
                if p.lead and p.max_deg > ((1 << ld) - (1 << (ls - 1))) and ls > 1:
                    ld = p.degree.bit_length()
                    return recurse(ts, ld >> 1, ld, p)
                return from_basis(ts, p)
            nxt = 1 << ls
            while nxt < (p.degree >> 1) + 1:
                nxt <<= 1
            q, r = split(p, nxt)
            level = C[nxt].level - 1
            if q.max_deg >= 1 << (ld - 1) and q.lead:
                level += 1
            qi = float(Q[level])
            res = recurse(ts * qi / C[nxt].scale, ls, ld, q)
            tmp = recurse(ts, ls, ld, r)
            if res.level > tmp.level:
                while res.level != tmp.level + 1:
                    res = self.drop_level(res, 1)
            res = self.mul_relin(res, C[nxt], rlk)
            if res.level > tmp.level:
                res = self.add_matched(self.rescale(res, eval_scale), tmp)
            else:
                res = self.rescale(self.add_matched(res, tmp), eval_scale)
            return res

        return recurse(target_scale, log_split, log_degree, P(list(coeffs), deg, True))

    RELU_COEFFS = ([0.0, 10.8541842577442, 0.0, -62.2833925211098, 0.0, 114.369227820443, 0.0, -62.8023496973074],
                   [0.0, 4.13976170985111, 0.0, -5.84997640211679, 0.0, 2.94376255659280, 0.0, -0.454530437460152],
                   [0.0, 3.29956739043733, 0.0, -7.84227260291355, 0.0, 12.8907764115564, 0.0, -12.4917112584486, 0.0,
                    6.94167991428074, 0.0, -2.04298067399942, 0.0, 0.246407138926031])

    def eval_relu(self, ct, alpha, rlk, eval_scale):
        """evalReLU (conv.go:435-480): three EvaluatePoly (minimax sign composition), AddConstNew, DropLevel,
        Mul + Relinearize.  Returns the level-(L-10) ciphertext at scale ct.scale * eval_scale (not rescaled)."""
        aconst, bconst = (alpha + 1) / 2.0, (1 - alpha) / 2.0
        s = ct
        for k, co in enumerate(self.RELU_COEFFS):
            s = self.evaluate_poly(s, [c * bconst for c in co] if k == 2 else co, eval_scale, rlk, eval_scale)
        s = self.add_const(s, aconst)
        x = self.drop_level(ct, ct.level - s.level)
        return self.mul_relin(s, x, rlk)

    def add_pt(self, ct, pt):
        lv = min(ct.level, pt.shape[0] - 1)
        x0, p = np.ascontiguousarray(ct.c0[:lv + 1]), np.ascontiguousarray(pt[:lv + 1])
        o0 = np.empty_like(x0)
        self.L.orc_add(self.h, lv, _p(x0), _p(p), _p(o0))
        return Ct(o0, ct.c1[:lv + 1].copy(), ct.scale)

    def keyswitch(self, c1, swk):
        c1 = np.ascontiguousarray(c1)
        lv = c1.shape[0] - 1
        d0, d1 = np.empty_like(c1), np.empty_like(c1)
        self.L.orc_keyswitch(self.h, lv, _p(c1), _p(swk), _p(d0), _p(d1))
        return d0, d1

    def decompose_digit(self, c1, d):
        """digit d of DecomposeSingleNTT for c1 [(level+1)][N] (NTT domain) -> (dQ, dP), NTT domain"""
        c1 = np.ascontiguousarray(c1)
        lv = c1.shape[0] - 1
        dQ = np.empty_like(c1)
        dP = np.empty((len(self.P), self.N), dtype=np.uint64)
        self.L.orc_decompose_digit(self.h, lv, d, _p(c1), _p(dQ), _p(dP))
        return dQ, dP

    def moddown(self, accQ, accP):
        accQ, accP = np.ascontiguousarray(accQ), np.ascontiguousarray(accP)
        lv = accQ.shape[0] - 1
        out = np.empty_like(accQ)
        self.L.orc_moddown(self.h, lv, _p(accQ), _p(accP), _p(out))
        return out

    def rotate_gal(self, ct, galEl, swk):
        o0, o1 = np.empty_like(ct.c0), np.empty_like(ct.c1)
        self.L.orc_rotate_gal(self.h, ct.level, _p(ct.c0), _p(ct.c1), galEl, _p(swk), _p(o0), _p(o1))
        return Ct(o0, o1, ct.scale)

    def rotate(self, ct, k, swk):
        return self.rotate_gal(ct, self.galois_for_rotation(k), swk)

    # ---- compositions ----
    def conv_then_pack(self, ct_in, pt_ker, pt_scale, norm, out_scale, pt_idx, swks, pt_bias=None,
                       nthreads=1):
        """evalConv_BN hot interval (eval.go:250-260).  swks: dict galEl-index j -> key array
        for galEl 2^(j+1)+1.  Returns (Ct, t_mult, t_pack)."""
        B = pt_ker.shape[0]
        arr = (u64p * self.logN)()
        keep = []
        for j in range(self.logN):
            if swks.get(j) is not None:
                k = np.ascontiguousarray(swks[j])
                keep.append(k)
                arr[j] = _p(k)
        o0, o1 = np.empty(self.N, dtype=np.uint64), np.empty(self.N, dtype=np.uint64)
        s, tm, tp = C.c_double(), C.c_double(), C.c_double()
        pt_ker = np.ascontiguousarray(pt_ker)
        pt_idx = np.ascontiguousarray(pt_idx)
        rc = self.L.orc_conv_then_pack(self.h, _p(ct_in.c0), _p(ct_in.c1), ct_in.scale, _p(pt_ker), pt_scale,
                                       B, norm, out_scale, _p(pt_idx), arr,
                                       _p(pt_bias) if pt_bias is not None else None, _p(o0), _p(o1),
                                       C.byref(s), nthreads, C.byref(tm), C.byref(tp))
        if rc != 0:
            raise RuntimeError("LV or scale after conv then pack, inconsistent")
        return Ct(o0[None, :], o1[None, :], s.value), tm.value, tp.value

    def conv_bl(self, ct_in, in_wid, ker_wid, rot_step, pt_taps, pt_scale, keys, pt_bias=None):
        """evalConv_BN_BL_test's timed interval (eval.go:108-131): preConv_BL (conv.go:120-143,
        RotateHoisted == the same rotations one by one, bit for bit) + postConv_BL (conv.go:146-178)
        + RotateNew/Add (eval.go:118-125) + bias (eval.go:130).  pt_taps: [rot_iters][k^2][L][N];
        keys: rotation k -> switching key."""
        h = ker_wid // 2
        rots = [i * in_wid + j for i in range(-h, h + 1) for j in range(-h, h + 1)]
        ct_rots = [ct_in if r == 0 else self.rotate(ct_in, r, keys[r]) for r in rots]
        res = None
        for i, taps in enumerate(pt_taps):
            tmp = None
            for t, pt in enumerate(taps):
                m = self.mul_pt(ct_rots[t], pt, pt_scale)
                tmp = m if tmp is None else self.add(tmp, m)
            if i == 0:
                res = tmp
            else:
                res = self.add(res, self.rotate(tmp, i * rot_step, keys[i * rot_step]))
        if pt_bias is not None:
            res = self.add_pt(res, pt_bias)
        return res

    def ext_ctxt(self, ct, r_idx, pt_scale, keys, min_scale=None):
        """ext_ctxt (conv.go:347-371): sum_rot RotateNew(MulNew(input, pt_rot), rot) [+ Rescale]."""
        res = None
        for rot, pt in r_idx.items():
            t = self.rotate(self.mul_pt(ct, pt, pt_scale), rot, keys[rot])
            res = t if res is None else self.add(res, t)
        return self.rescale(res, min_scale) if min_scale is not None else res

    def ext_double_ctxt(self, ct, m_idx, r_idx, pt_scale, keys, min_scale):
        """ext_double_ctxt (conv.go:374-414) / bsgs_ctxt (conv.go:303-344, which does not rescale: min_scale None):
        mid = sum_rot RotateNew(MulNew(input, pt_rot), rot) over m_idx, result = the same over r_idx applied to mid."""
        mid = self.ext_ctxt(ct, m_idx, pt_scale, keys)
        return self.ext_ctxt(mid, r_idx, pt_scale, keys, min_scale)

    def post_conv_bl(self, ct_in_rots, pts, pt_scale):
        """postConv_BL (conv.go:146-178): sum_tap MulNew(ct_in_rots[tap], pl_tap)."""
        res = None
        for c, pt in zip(ct_in_rots, pts):
            t = self.mul_pt(c, pt, pt_scale)
            res = t if res is None else self.add(res, t)
        return res

    def keep_ctxt(self, ct, mask, pt_scale, min_scale):
        """keep_ctxt (conv.go:417-431)."""
        return self.rescale(self.mul_pt(ct, mask, pt_scale), min_scale)

    @staticmethod
    def conv_bn_relu(op, o, ct_input, *, pt_ker, pt_bias, pt_scale, norm, out_scale, pt_idx, pack_keys, pow, alpha, iter, btp,
                     keys, key_conj, rlk, stoc_mats, min_scale, keep_mask=None, keep_scale=None, r_idx=None, m_idx=None, mask_scale=None,
                     pt_pre=None, pt_shift2=None, pt_post=None):
        """evalConv_BNRelu_new (eval.go:272-575) composed from the pinned pieces, on the two evaluators (op: pack, o: main):
        [x^offset] -> evalConv_BN once or twice (+ shift, Add) -> [x^offset] -> Scale *= 2^pow -> BootstrappConv_CtoS ->
        evalReLU + MulByPow2 per half -> keep_ctxt / ext_ctxt / ext_double_ctxt per half -> BootstrappConv_StoC -> Rescale."""
        cin = op.mul_pt(ct_input, pt_pre, 1.0) if pt_pre is not None else ct_input
        convs = [op.conv_then_pack(cin, pt_ker[k], pt_scale, norm[k], out_scale, pt_idx, pack_keys, pt_bias[k])[0]
                 for k in range(len(pt_ker))]
        if len(convs) == 2:
            c1 = op.mul_pt(convs[1], pt_shift2, 1.0) if pt_shift2 is not None else convs[1]
            ct_conv = op.add(convs[0], c1)
        else:
            ct_conv = convs[0]
        if pt_post is not None:
            ct_conv = op.mul_pt(ct_conv, pt_post, 1.0)
        ct_conv = Ct(ct_conv.c0, ct_conv.c1, ct_conv.scale * 2.0 ** pow)
        b0, b1, _ = o.bootstrapp_conv_ctos(ct_conv, btp, keys, key_conj, rlk)
        halves = [b0, b1][:iter]
        kept = []
        for ul, h in enumerate(halves):
            if h is None:
                kept.append(None)
                continue
            x = o.mul_by_pow2(o.eval_relu(h, alpha, rlk, min_scale), int(pow))
            if keep_mask is not None:
                kept.append(o.keep_ctxt(x, keep_mask[ul], keep_scale, min_scale))
            elif m_idx is not None:
                kept.append(o.ext_double_ctxt(x, m_idx[ul], r_idx[ul], mask_scale, keys, min_scale))
            else:
                kept.append(o.ext_ctxt(x, r_idx[ul], mask_scale, keys, min_scale))
        res = o.slots_to_coeffs(kept[0], kept[1] if iter == 2 else None, stoc_mats, keys)
        return o.rescale(res, min_scale)

    def monomial_pts(self):
        """pl_idx[i] = NTT(X^(2^i)) at level 0, scale 1 (conv.go:241-254)."""
        out = np.empty((self.logN, self.N), dtype=np.uint64)
        for i in range(self.logN):
            m = np.zeros(self.N, dtype=np.uint64)
            m[1 << i] = 1
            out[i] = self.ntt(m, 0)
        return out

    # ---- semantic helpers ----
    def gen_secret(self, seed, h=192):
        sQ = np.empty((len(self.Q), self.N), dtype=np.uint64)
        sP = np.empty((max(len(self.P), 1), self.N), dtype=np.uint64)
        self.L.orc_gen_secret(self.h, seed, h, _p(sQ), _p(sP))
        return sQ, sP

    def gen_rotkey(self, seed, galEl, sQ, sP):
        swk = np.empty((self.beta_full, 2, len(self.Q) + len(self.P), self.N), dtype=np.uint64)
        self.L.orc_gen_rotkey(self.h, seed, galEl, _p(sQ), _p(sP), _p(swk))
        return swk

    def encrypt(self, seed, m, sQ):
        m = np.ascontiguousarray(m)
        lv = m.shape[0] - 1
        c0, c1 = np.empty_like(m), np.empty_like(m)
        self.L.orc_encrypt(self.h, seed, lv, _p(m), _p(sQ), _p(c0), _p(c1))
        return c0, c1

    def decrypt(self, ct, sQ):
        m = np.empty_like(ct.c0)
        self.L.orc_decrypt(self.h, ct.level, _p(ct.c0), _p(ct.c1), _p(sQ), _p(m))
        return m
