"""The reference binary as an interpreted machine: x86emu.Emu + the few Go-runtime services its pure
arithmetic code needs (allocation, memmove, maps, math/big), all provided in Python.

Used only by tests/golden/make_ref_eval_vectors.py to produce golden vectors from the reference's OWN
compiled evaluator code (Lattigo fork ring/rlwe/ckks packages and main.conv_then_pack) without executing
the prebuilt binary and without a Go toolchain.  Everything under github.com/dwkim606/test_lattigo/, main.
and math. is interpreted from its disassembly; calls into the Go runtime are answered by the hooks below;
anything else raises.  Go 1.16 stack ABI: arguments then results above the return address.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

from x86emu import Emu, M64, b2f, f2b  # noqa: E402

BIN = "/root/reference/test_run"
LAT = "github.com/dwkim606/test_lattigo/"
RING = LAT + "ring."
RLWE = LAT + "rlwe."
CKKS = LAT + "ckks."


class GoPanic(Exception):
    pass


def _factors(n):
    out, p = [], 2
    while p * p <= n:
        if n % p == 0:
            out.append(p)
            while n % p == 0:
                n //= p
        p += 1 if p == 2 else 2
    if n > 1:
        out.append(n)
    return out


class Machine(Emu):
    auto_prefixes = (LAT, "main.", "math.", "math/bits.", "runtime.duffcopy", "runtime.duffzero",
                     # pure float helpers of the Go runtime (complex division)
                     "runtime.complex128div", "runtime.inf2one", "runtime.isNaN", "runtime.isInf", "runtime.isFinite",
                     "runtime.abs", "runtime.copysign", "runtime.float64bits", "runtime.float64frombits")

    def __init__(self):
        super().__init__(BIN)
        self.maps = {}      # handle -> {key: slot address}
        self.mapval = {}    # handle -> value size
        self.zero = self.alloc(1024)
        H = self.hook
        H("runtime.newobject", self.h_newobject)
        H("runtime.makeslice", self.h_makeslice)
        H("runtime.memmove", self.h_memmove)
        H("runtime.memclrNoHeapPointers", self.h_memclr)
        H("runtime.typedmemmove", self.h_typedmemmove)
        H("runtime.convT64", self.h_convT64)
        H("runtime.mapaccess1_fast64", lambda em: self.h_mapaccess(1))
        H("runtime.mapaccess2_fast64", lambda em: self.h_mapaccess(2))
        H("runtime.mapassign_fast64", self.h_mapassign)
        H("runtime.gopanic", self.h_panic)
        for n in ("panicIndex", "panicIndexU", "panicSliceAcap", "panicSliceAlen", "panicSliceB", "panicdivide",
                  "panicshift", "panicSliceAcapU", "panicSliceAlenU", "panicSliceBU", "panicSlice3Alen", "panicSlice3C"):
            if "runtime." + n in self.sym:
                H("runtime." + n, self._mk_panic(n))
        for n in ("time.Now", "time.Since", "fmt.Fprintln", "fmt.Println", "fmt.Printf", "fmt.Fprintf"):
            if n in self.sym:
                H(n, lambda em: None)
        # math/big (Int = {neg bool; abs nat{ptr,len,cap}})
        H("math/big.(*Int).SetInt64", self.h_big_setint64)
        H("math/big.nat.setUint64", self.h_nat_set)
        for n, fn in (("Mul", lambda a, b: a * b), ("Mod", lambda a, b: a % b), ("Quo", lambda a, b: abs(a) // abs(b) * (1 if (a < 0) == (b < 0) else -1)),
                      ("ModInverse", lambda a, b: pow(a, -1, b)), ("Add", lambda a, b: a + b), ("Sub", lambda a, b: a - b)):
            if "math/big.(*Int)." + n in self.sym:
                H("math/big.(*Int)." + n, self._mk_bigop(fn))
        # math/big.Float at its default 53-bit precision, ToNearestEven == IEEE double arithmetic
        # (ckks.scaleUpExact: NewFloat(x); Add(.,0.5); Int).  The double lives in the struct's mant.ptr slot.
        H("math/big.(*Float).SetFloat64", self.h_flt_set)
        H("math/big.(*Float).Add", self.h_flt_add)
        H("math/big.(*Float).Int", self.h_flt_int)
        H("runtime.ifaceeq", self.h_ifaceeq)
        H("runtime.makemap_small", lambda em: self.wq(self.r[4], self.new_map(None)))
        H("runtime.makemap", lambda em: self.wq(self.r[4] + 24, self.new_map(None)))
        H("runtime.mapiterinit", self.h_mapiterinit)
        H("runtime.fastrand", self.h_fastrand)
        H("runtime.growslice", self.h_growslice)
        H("fmt.Errorf", self.h_errorf)
        H("runtime.typedslicecopy", self.h_typedslicecopy)
        H("runtime.mapiternext", self.h_mapiternext)
        # pure number theory whose Go implementation needs math/big or a prime sieve
        H(RING + "getFactors", self.h_getfactors)
        H(RING + "BRedParams", self.h_bredparams)
        H(RING + "IsPrime", lambda em: em.wb(em.r[4] + 8, 1))
        H(LAT + "utils.AllDistinct", lambda em: em.wb(em.r[4] + 24, 1))

    # ---- runtime services ----
    def h_newobject(self, em):
        sp = self.r[4]
        self.wq(sp + 8, self.alloc(max(8, self.load_const(self.rq(sp)))))

    def h_makeslice(self, em):
        sp = self.r[4]
        self.wq(sp + 24, self.alloc(max(8, self.load_const(self.rq(sp)) * self.rq(sp + 16))))

    def copy_bytes(self, dst, src, n):
        if n == 0 or dst == src:
            return
        if (dst | src | n) & 7 == 0:
            vals = [self.mem.get(src + 8 * i, 0) for i in range(n // 8)]
            for i, v in enumerate(vals):
                self.mem[dst + 8 * i] = v
        else:
            vals = [self.rd(src + i, 1) for i in range(n)]
            for i, v in enumerate(vals):
                self.wb(dst + i, v)

    def h_memmove(self, em):
        sp = self.r[4]
        self.copy_bytes(self.rq(sp), self.rq(sp + 8), self.rq(sp + 16))

    def h_typedmemmove(self, em):
        sp = self.r[4]
        self.copy_bytes(self.rq(sp + 8), self.rq(sp + 16), self.load_const(self.rq(sp)))

    def h_memclr(self, em):
        sp = self.r[4]
        p, n = self.rq(sp), self.rq(sp + 8)
        assert (p | n) & 7 == 0
        for i in range(n // 8):
            self.mem[p + 8 * i] = 0

    def h_convT64(self, em):
        sp = self.r[4]
        a = self.alloc(8)
        self.wq(a, self.rq(sp))
        self.wq(sp + 8, a)

    def _mk_panic(self, n):
        def h(em):
            raise GoPanic("runtime." + n)
        return h

    def h_panic(self, em):
        sp = self.r[4]
        typ, data = self.rq(sp), self.rq(sp + 8)
        msg = ""
        try:  # string or error value: best effort decode of a Go string header
            p, ln = self.rq(data), self.rq(data + 8)
            if 0 < ln < 200:
                msg = bytes(self.byte_at(p + i) for i in range(ln)).decode("utf8", "replace")
        except Exception:
            pass
        raise GoPanic("gopanic: " + msg)

    def byte_at(self, a):
        if (a & ~7) in self.mem:
            return self.rd(a, 1)
        return (self.load_const(a & ~7) >> ((a & 7) * 8)) & 0xff

    # maps with uint64 keys (rtks.Keys, permuteNTTIndex): python dicts behind an opaque handle
    def new_map(self, valsize):
        h = self.alloc(64)
        self.maps[h] = {}
        self.mapval[h] = valsize
        return h

    def map_put(self, h, key, words):
        slot = self.alloc(self.mapval[h])
        self.write_u64s(slot, words)
        self.maps[h][key] = slot

    def _mapid(self, h):
        """maps the compiled code makes on its own stack (hmap zeroed + hash0 = fastrand()) are told apart
        from stale ones at the same address by the unique hash0 our fastrand hook hands out"""
        return (h, self.rd(h + 12, 4)) if h and h not in self.maps else h

    def h_growslice(self, em):
        sp = self.r[4]
        esz = self.load_const(self.rq(sp))
        ptr, ln, cap, want = self.rq(sp + 8), self.rq(sp + 16), self.rq(sp + 24), self.rq(sp + 32)
        ncap = max(want, 2 * cap)
        a = self.alloc(max(8, esz * ncap))
        self.copy_bytes(a, ptr, esz * ln)
        self.write_u64s(sp + 40, [a, ln, ncap])

    def h_typedslicecopy(self, em):
        sp = self.r[4]
        esz = self.load_const(self.rq(sp))
        n = min(self.rq(sp + 16), self.rq(sp + 32))
        self.copy_bytes(self.rq(sp + 8), self.rq(sp + 24), esz * n)
        self.wq(sp + 40, n)

    def h_errorf(self, em):
        """fmt.Errorf(format, args...) -> a non-nil error whose data word points at the (unformatted) format string"""
        sp = self.r[4]
        data = self.alloc(16)
        self.write_u64s(data, [self.rq(sp), self.rq(sp + 8)])
        self.write_u64s(sp + 40, [self.alloc(32), data])

    def h_fastrand(self, em):
        self.rand_ctr = getattr(self, "rand_ctr", 0) + 1
        self.wr(self.r[4], 4, self.rand_ctr)

    def h_mapaccess(self, nres):
        sp = self.r[4]
        h, key = self._mapid(self.rq(sp + 8)), self.rq(sp + 16)
        slot = self.maps[h].get(key) if h in self.maps else None
        self.wq(sp + 24, slot if slot is not None else self.zero)
        if nres == 2:
            self.wb(sp + 32, int(slot is not None))

    def h_mapassign(self, em):
        sp = self.r[4]
        t, h0, key = self.rq(sp), self.rq(sp + 8), self.rq(sp + 16)
        h = self._mapid(h0)
        if h not in self.maps:
            self.maps[h], self.mapval[h] = {}, None
        if key not in self.maps[h]:
            self.wq(h0, self.rq(h0) + 1)     # hmap.count (len(m) is read inline)
        if self.mapval[h] is None:     # maptype.elem (*_type at +56) -> _type.size (first word)
            self.mapval[h] = self.load_const(self.load_const(t + 56))
        if key not in self.maps[h]:
            self.maps[h][key] = self.alloc(self.mapval[h])
        self.wq(sp + 24, self.maps[h][key])

    # ---- math/big ----
    def big_read(self, a):
        ptr, n = self.rq(a + 8), self.rq(a + 16)
        v = 0
        for i in range(n):
            v |= self.rq(ptr + 8 * i) << (64 * i)
        return -v if self.rd(a, 1) else v

    def nat_alloc(self, v):
        ws = []
        while v:
            ws.append(v & M64)
            v >>= 64
        p = self.alloc(8 * max(1, len(ws)))
        self.write_u64s(p, ws)
        return p, len(ws)

    def big_write(self, a, v):
        self.wb(a, 1 if v < 0 else 0)
        p, n = self.nat_alloc(abs(v))
        self.wq(a + 8, p)
        self.wq(a + 16, n)
        self.wq(a + 24, max(1, n))

    def h_big_setint64(self, em):
        sp = self.r[4]
        z, x = self.rq(sp), self.rq(sp + 8)
        self.big_write(z, x if x < (1 << 63) else x - (1 << 64))
        self.wq(sp + 16, z)

    def h_nat_set(self, em):
        sp = self.r[4]
        p, n = self.nat_alloc(self.rq(sp + 24))
        self.wq(sp + 32, p)
        self.wq(sp + 40, n)
        self.wq(sp + 48, max(1, n))

    def _mk_bigop(self, fn):
        def h(em):
            sp = self.r[4]
            z, x, y = self.rq(sp), self.rq(sp + 8), self.rq(sp + 16)
            self.big_write(z, fn(self.big_read(x), self.big_read(y)))
            self.wq(sp + 24, z)
        return h

    def h_flt_set(self, em):
        sp = self.r[4]
        z = self.rq(sp)
        self.wq(z + 8, self.rq(sp + 8))
        self.wq(sp + 16, z)

    def h_flt_add(self, em):
        sp = self.r[4]
        z, x, y = self.rq(sp), self.rq(sp + 8), self.rq(sp + 16)
        self.wq(z + 8, f2b(b2f(self.rq(x + 8)) + b2f(self.rq(y + 8))))
        self.wq(sp + 24, z)

    def h_flt_int(self, em):
        sp = self.r[4]
        x, z = self.rq(sp), self.rq(sp + 8)
        if z == 0:
            z = self.alloc(32)
        f = b2f(self.rq(x + 8))
        self.big_write(z, int(f))
        self.wq(sp + 16, z)
        self.wb(sp + 24, 0 if f == int(f) else (0xff if f > 0 else 1))

    def h_ifaceeq(self, em):
        sp = self.r[4]
        self.wb(sp + 24, int(self.rq(sp + 8) == self.rq(sp + 16)))

    def h_mapiterinit(self, em):
        sp = self.r[4]
        h, it = self._mapid(self.rq(sp + 8)), self.rq(sp + 16)
        self.iters = getattr(self, "iters", {})
        self.iters[it] = iter(sorted(self.maps[h].items())) if h in self.maps else iter(())
        self._iter_step(it)

    def h_mapiternext(self, em):
        self._iter_step(self.rq(self.r[4]))

    def _iter_step(self, it):
        try:
            k, slot = next(self.iters[it])
            kp = self.alloc(8)
            self.wq(kp, k)
            self.wq(it, kp)
            self.wq(it + 8, slot)
        except StopIteration:
            self.wq(it, 0)
            self.wq(it + 8, 0)

    def h_getfactors(self, em):
        sp = self.r[4]
        fs = _factors(self.rq(sp))
        a = self.alloc(8 * len(fs))
        self.write_u64s(a, fs)
        self.write_u64s(sp + 8, [a, len(fs), len(fs)])

    def h_bredparams(self, em):
        sp = self.r[4]
        b = (1 << 128) // self.rq(sp)
        a = self.alloc(16)
        self.write_u64s(a, [b >> 64, b & M64])
        self.write_u64s(sp + 8, [a, 2, 2])

    # ---- Go values ----
    def slice_u64(self, vals):
        """[]uint64 -> (ptr, len, cap)"""
        a = self.alloc(8 * max(1, len(vals)))
        self.write_u64s(a, vals)
        return [a, len(vals), len(vals)]

    def read_slice_u64(self, hdr):
        return self.read_u64s(self.rq(hdr), self.rq(hdr + 8))

    def new_poly(self, limbs):
        """*ring.Poly {Coeffs [][]uint64; IsNTT, IsMForm bool} from a list of limb lists"""
        rows = self.alloc(24 * max(1, len(limbs)))
        for i, l in enumerate(limbs):
            self.write_u64s(rows + 24 * i, self.slice_u64(l))
        p = self.alloc(32)
        self.write_u64s(p, [rows, len(limbs), len(limbs), 0])
        return p

    def read_poly(self, p, nlimbs=None):
        rows, n = self.rq(p), self.rq(p + 8)
        n = n if nlimbs is None else nlimbs
        return [self.read_slice_u64(rows + 24 * i) for i in range(n)]

    def new_ring(self, N, moduli):
        res = self.call(RING + "NewRing", [N] + self.slice_u64(moduli) + [0, 0, 0])
        if res[5] or res[6]:
            raise GoPanic("NewRing returned an error")
        return res[4]

    # ckks.Ciphertext{*rlwe.Ciphertext{Value []*ring.Poly}; Scale float64}; polys are NTT-domain (IsNTT set)
    def new_ct(self, polys, scale):
        ptrs = []
        for limbs in polys:
            p = self.new_poly(limbs)
            self.wb(p + 24, 1)
            ptrs.append(p)
        inner = self.alloc(24)
        self.write_u64s(inner, self.slice_u64(ptrs))
        ct = self.alloc(16)
        self.write_u64s(ct, [inner, f2b(scale)])
        return ct

    def read_ct(self, ct):
        inner = self.rq(ct)
        ptrs = self.read_slice_u64(inner)
        return [self.read_poly(p) for p in ptrs], b2f(self.rq(ct + 8))

    # ckks.Plaintext{*rlwe.Plaintext{Value *ring.Poly}; Scale float64}
    def new_pt(self, limbs, scale):
        p = self.new_poly(limbs)
        self.wb(p + 24, 1)
        inner = self.alloc(8)
        self.wq(inner, p)
        pt = self.alloc(16)
        self.write_u64s(pt, [inner, f2b(scale)])
        return pt

    def new_swk(self, swk):
        """*rlwe.SwitchingKey{Value [][2]*ring.Poly} from an array [beta][2][nQ+nP][N] (NTT + Montgomery form)"""
        beta = len(swk)
        val = self.alloc(16 * beta)
        for d in range(beta):
            for k in range(2):
                pol = self.new_poly([[int(v) for v in limb] for limb in swk[d][k]])
                self.wb(pol + 24, 1)
                self.wb(pol + 25, 1)
                self.wq(val + 16 * d + 8 * k, pol)
        key = self.alloc(24)
        self.write_u64s(key, [val, beta, beta])
        return key

    def new_evaluator(self, logN, Q, P, scale, keys, rlk=None, log_slots=None):
        """ckks.NewEvaluator(params, EvaluationKey{Rlk: rlk, Rtks: keys}) run by the reference code itself.
        keys: {galEl: array [beta][2][nQ+nP][N]}; rlk: one such array (RelinearizationKey{Keys []*SwitchingKey})
        or None.  Returns (params words, Evaluator iface words)."""
        rQ, rP = self.new_ring(1 << logN, Q), self.new_ring(1 << logN, P)
        params = [logN] + self.slice_u64(Q) + self.slice_u64(P) + [f2b(3.2), rQ, rP, 0, logN - 1 if log_slots is None else log_slots, f2b(scale)]
        kmap = self.new_map(8)
        for gal, swk in keys.items():
            self.map_put(kmap, gal, [self.new_swk(swk)])
        rtks = self.alloc(8)
        self.wq(rtks, kmap)
        rlkp = 0
        if rlk is not None:
            rlkp = self.alloc(24)
            self.write_u64s(rlkp, self.slice_u64([self.new_swk(rlk)]))
        res = self.call(CKKS + "NewEvaluator", params + [rlkp, rtks, 0, 0])
        return params, res[-2:]
