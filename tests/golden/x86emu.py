"""A small x86-64 interpreter over `objdump -d` text, used ONLY to generate golden vectors from the
reference's own compiled code (tests/golden/make_ref_vectors.py).

Why: the reference's arithmetic lives in an un-vendored Go module and there is no Go toolchain, but
the prebuilt binary /root/reference/test_run carries the compiled routines.  The binary is never
executed: the pure leaf routines (ring.NTTLazy, ring.InvNTT, ring.reconstructRNS, ...) are
*interpreted* here, instruction by instruction, from their disassembly, on a private register file
and a private flat memory.  No system call, no I/O, no native execution; a call to anything that is
not one of the whitelisted pure routines raises.

Only the instruction forms those routines use are implemented (see _exec); anything else raises.
Go 1.16 stack ABI: arguments and results live on the caller's stack above the return address.
"""
import re
import struct
import subprocess

M64 = (1 << 64) - 1

_R64 = ["rax", "rcx", "rdx", "rbx", "rsp", "rbp", "rsi", "rdi"] + ["r%d" % i for i in range(8, 16)]
_R32 = ["eax", "ecx", "edx", "ebx", "esp", "ebp", "esi", "edi"] + ["r%dd" % i for i in range(8, 16)]
_R16 = ["ax", "cx", "dx", "bx", "sp", "bp", "si", "di"] + ["r%dw" % i for i in range(8, 16)]
_R8 = ["al", "cl", "dl", "bl", "spl", "bpl", "sil", "dil"] + ["r%db" % i for i in range(8, 16)]
REGS = {}
for _i in range(16):
    REGS[_R64[_i]] = (_i, 8)
    REGS[_R32[_i]] = (_i, 4)
    REGS[_R16[_i]] = (_i, 2)
    REGS[_R8[_i]] = (_i, 1)

_MEM_RE = re.compile(r"^(?:%(\w+):)?(-?0x[0-9a-f]+|-?\d+)?(?:\((%\w+)?(?:,(%\w+))?(?:,(\d+))?\))?$")


def _split_ops(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def _parse_op(tok, comment_addr):
    if tok.startswith("$"):
        return ("i", int(tok[1:], 16) if "0x" in tok else int(tok[1:]))
    if tok.startswith("%") and ":" not in tok:
        name = tok[1:]
        if name.startswith("xmm"):
            return ("x", int(name[3:]))
        return ("r",) + REGS[name]
    m = _MEM_RE.match(tok)
    if not m:
        raise ValueError("operand " + tok)
    seg, disp, base, index, scale = m.groups()
    d = int(disp, 16) if disp and "0x" in disp else (int(disp) if disp else 0)
    if base == "%rip":
        return ("c", comment_addr)
    b = REGS[base[1:]][0] if base else None
    ix = REGS[index[1:]][0] if index else None
    return ("m", seg, d, b, ix, int(scale) if scale else 1)


class Insn:
    __slots__ = ("addr", "mn", "ops", "next", "target")


def disassemble(path, start, size):
    txt = subprocess.run(["objdump", "-d", "--no-show-raw-insn", "--start-address=0x%x" % start,
                          "--stop-address=0x%x" % (start + size), path], capture_output=True, text=True, check=True).stdout
    insns = []
    for line in txt.splitlines():
        m = re.match(r"^\s+([0-9a-f]+):\t(\S+)\s*(.*)$", line)
        if not m:
            continue
        ins = Insn()
        ins.addr = int(m.group(1), 16)
        ins.mn = m.group(2)
        rest = m.group(3)
        caddr = None
        if "#" in rest:
            rest, com = rest.split("#", 1)
            caddr = int(com.split()[0], 16)
        rest = re.sub(r"<.*?>", "", rest).strip()
        ins.target = None
        if (ins.mn.startswith("j") or ins.mn == "call") and rest.startswith("*"):
            ins.ops = [_parse_op(rest[1:], caddr)]     # indirect: target read at run time
        elif ins.mn.startswith("j") or ins.mn == "call":
            ins.target = int(rest.split()[0], 16)
            ins.ops = []
        else:
            ins.ops = [_parse_op(t, caddr) for t in _split_ops(rest)] if rest else []
        insns.append(ins)
    for a, b in zip(insns, insns[1:]):
        a.next = b.addr
    insns[-1].next = None
    return insns


def symbols(path):
    out = subprocess.run(["nm", "-S", path], capture_output=True, text=True, check=True).stdout
    tab = {}
    for line in out.splitlines():
        f = line.split()
        if len(f) == 4:
            tab[f[3]] = (int(f[0], 16), int(f[1], 16))
    return tab


def _sx(v, size):
    bits = size * 8
    v &= (1 << bits) - 1
    return v - (1 << bits) if v >> (bits - 1) else v


def f2b(x):
    return struct.unpack("<Q", struct.pack("<d", x))[0]


def b2f(b):
    return struct.unpack("<d", struct.pack("<Q", b & M64))[0]


class Emu:
    def __init__(self, path):
        self.path = path
        self.sym = symbols(path)
        self.code = {}      # addr -> Insn
        self.allowed = set()
        self.mem = {}       # qword-aligned address -> 64-bit int
        self.r = [0] * 16
        self.x = [0] * 16   # 128-bit ints
        self.cf = self.zf = self.sf = self.of = self.pf = 0
        self.rodata = {}
        self.steps = 0
        self.G = 0x7f0000000000  # fake g struct: stackguard0 = 0
        self.wq(self.G + 0x10, 0)
        self.hooks = {}     # call target -> python function(emu): reads args at [rsp], writes results after them
        self.tracers = {}   # call target -> python observer(emu), called before the call is made
        self.heap = 0x600000000000

    def alloc(self, nbytes):
        a = self.heap
        self.heap += (nbytes + 63) & ~63
        return a

    def hook(self, name, fn):
        self.hooks[self.sym[name][0]] = fn

    # ---- loading ----
    def load(self, name):
        start, size = self.sym[name]
        for ins in disassemble(self.path, start, size):
            self.code[ins.addr] = ins
        self.allowed.add(start)
        return start

    def name_of(self, addr):
        """symbol containing addr (for messages and the autoload whitelist)"""
        if not hasattr(self, "_ranges"):
            self._ranges = sorted((a, a + sz, n) for n, (a, sz) in self.sym.items() if sz)
        import bisect
        i = bisect.bisect_right(self._ranges, (addr, 1 << 62, "")) - 1
        if i >= 0 and self._ranges[i][0] <= addr < self._ranges[i][1]:
            return self._ranges[i][2]
        return None

    auto_prefixes = ()

    def autoload(self, addr):
        """load the routine containing addr if its symbol is in a whitelisted (pure-Go) package"""
        name = self.name_of(addr)
        if name is None or not name.startswith(self.auto_prefixes):
            return False
        self.load(name)
        return addr in self.code

    def load_const(self, addr):
        """8 bytes of the binary's read-only data at virtual address addr"""
        if addr in self.rodata:
            return self.rodata[addr]
        from elftools.elf.elffile import ELFFile
        with open(self.path, "rb") as f:
            elf = ELFFile(f)
            for seg in elf.iter_segments():
                if seg["p_type"] == "PT_LOAD" and seg["p_vaddr"] <= addr < seg["p_vaddr"] + seg["p_filesz"]:
                    f.seek(seg["p_offset"] + addr - seg["p_vaddr"])
                    v = struct.unpack("<Q", f.read(8))[0]
                    self.rodata[addr] = v
                    return v
                if seg["p_type"] == "PT_LOAD" and seg["p_vaddr"] <= addr < seg["p_vaddr"] + seg["p_memsz"]:
                    return 0  # .bss / .noptrbss: zero at load time (e.g. runtime.writeBarrier.enabled)
        raise ValueError("constant outside file at %x" % addr)

    # ---- memory ----
    def _image(self, a):
        """aligned qword not written by this run: the binary's own data (itabs, type descriptors) or zero"""
        if 0x400000 <= a < 0x1000000:
            try:
                return self.load_const(a)
            except ValueError:
                return 0
        return 0

    def rq(self, a):
        if a & 7 == 0:
            v = self.mem.get(a)
            return v if v is not None else self._image(a)
        lo = self.mem.get(a & ~7)
        hi = self.mem.get((a & ~7) + 8)
        lo = lo if lo is not None else self._image(a & ~7)
        hi = hi if hi is not None else self._image((a & ~7) + 8)
        sh = (a & 7) * 8
        return ((lo >> sh) | (hi << (64 - sh))) & M64

    def wq(self, a, v):
        v &= M64
        if a & 7 == 0:
            self.mem[a] = v
            return
        for i in range(8):
            self.wb(a + i, (v >> (8 * i)) & 0xff)

    def wb(self, a, v):
        base, sh = a & ~7, (a & 7) * 8
        self.mem[base] = (self.mem.get(base, 0) & ~(0xff << sh) & M64) | ((v & 0xff) << sh)

    def rd(self, a, size):
        return self.rq(a) & ((1 << (size * 8)) - 1)

    def wr(self, a, size, v):
        if size == 8:
            self.wq(a, v)
        else:
            for i in range(size):
                self.wb(a + i, (v >> (8 * i)) & 0xff)

    def write_u64s(self, addr, vals):
        for i, v in enumerate(vals):
            self.mem[addr + 8 * i] = int(v) & M64

    def read_u64s(self, addr, n):
        return [self.mem.get(addr + 8 * i, 0) for i in range(n)]

    # ---- operands ----
    def ea(self, op):
        _, seg, disp, b, ix, sc = op
        a = disp
        if b is not None:
            a += self.r[b]
        if ix is not None:
            a += self.r[ix] * sc
        a &= M64
        if seg == "fs":
            if a == (M64 - 7):  # %fs:-8 -> g
                return ("G",)
            raise ValueError("fs access %x" % a)
        return a

    def get(self, op, size):
        k = op[0]
        if k == "r":
            return self.r[op[1]] & ((1 << (op[2] * 8)) - 1)
        if k == "i":
            return op[1] & ((1 << (size * 8)) - 1)
        if k == "c":
            if (op[1] & ~7) in self.mem:        # a global this run has written
                return self.rd(op[1], size)
            return (self.load_const(op[1] & ~7) >> ((op[1] & 7) * 8)) & ((1 << (size * 8)) - 1) if op[1] & 7 and size < 8 \
                else self.load_const(op[1]) & ((1 << (size * 8)) - 1)
        if k == "m":
            a = self.ea(op)
            if a == ("G",):
                return self.G
            return self.rd(a, size)
        raise ValueError(op)

    def put(self, op, size, v):
        if op[0] == "r":
            i, s = op[1], op[2]
            if s == 8:
                self.r[i] = v & M64
            elif s == 4:
                self.r[i] = v & 0xffffffff
            else:
                m = (1 << (s * 8)) - 1
                self.r[i] = (self.r[i] & ~m) | (v & m)
        elif op[0] == "c":
            if (op[1] & ~7) not in self.mem:
                self.mem[op[1] & ~7] = self.load_const(op[1] & ~7)
            self.wr(op[1], size, v)
        else:
            self.wr(self.ea(op), size, v)

    @staticmethod
    def opsize(ops, mn):
        for o in ops:
            if o[0] == "r":
                return o[2]
        return {"b": 1, "w": 2, "l": 4, "q": 8}.get(mn[-1], 8)

    def setflags_logic(self, res, size):
        bits = size * 8
        res &= (1 << bits) - 1
        self.zf = int(res == 0)
        self.sf = res >> (bits - 1)
        self.cf = self.of = 0

    def setflags_sub(self, a, b, size, borrow=0):
        bits = size * 8
        m = (1 << bits) - 1
        res = (a - b - borrow) & m
        self.cf = int(a < b + borrow)
        self.zf = int(res == 0)
        self.sf = res >> (bits - 1)
        sa, sb, sr = a >> (bits - 1), b >> (bits - 1), res >> (bits - 1)
        self.of = int(sa != sb and sr != sa)
        return res

    def setflags_add(self, a, b, size, carry=0):
        bits = size * 8
        m = (1 << bits) - 1
        full = a + b + carry
        res = full & m
        self.cf = int(full > m)
        self.zf = int(res == 0)
        self.sf = res >> (bits - 1)
        sa, sb, sr = a >> (bits - 1), b >> (bits - 1), res >> (bits - 1)
        self.of = int(sa == sb and sr != sa)
        return res

    def cond(self, cc):
        if cc in ("ae", "nb", "nc"): return not self.cf
        if cc in ("b", "c", "nae"): return bool(self.cf)
        if cc in ("be", "na"): return bool(self.cf or self.zf)
        if cc in ("a", "nbe"): return not (self.cf or self.zf)
        if cc in ("e", "z"): return bool(self.zf)
        if cc in ("ne", "nz"): return not self.zf
        if cc in ("l", "nge"): return self.sf != self.of
        if cc in ("le", "ng"): return bool(self.zf or self.sf != self.of)
        if cc in ("g", "nle"): return (not self.zf) and self.sf == self.of
        if cc in ("ge", "nl"): return self.sf == self.of
        if cc == "s": return bool(self.sf)
        if cc == "ns": return not self.sf
        if cc == "p": return bool(self.pf)
        if cc == "np": return not self.pf
        raise ValueError("cc " + cc)

    # ---- execution ----
    def call(self, name, stack_words, max_steps=400_000_000):
        """Call function `name` with the Go stack-ABI argument/result area `stack_words` (list of u64 placed
        right above the return address).  Returns the argument/result area after the call."""
        entry = self.sym[name][0]
        if entry not in self.code:
            self.load(name)
        sp = 0x7e0000000000
        self.r[4] = sp
        self.wq(sp, 0xdeadbeef)  # sentinel return address
        for i, w in enumerate(stack_words):
            self.wq(sp + 8 + 8 * i, w)
        pc = entry
        code = self.code
        while pc != 0xdeadbeef:
            ins = code.get(pc)
            if ins is None:
                if not self.autoload(pc):    # tail jump into another whitelisted routine
                    raise RuntimeError("pc %x (%s) outside loaded routines" % (pc, self.name_of(pc)))
                ins = code[pc]
            try:
                pc = self._exec(ins)
            except Exception as ex:
                raise RuntimeError("at %x in %s: %s %s: %r" % (ins.addr, self.name_of(ins.addr), ins.mn, ins.ops, ex)) from ex
            self.steps += 1
            if self.steps > max_steps:
                raise RuntimeError("step limit")
        return [self.rq(sp + 8 + 8 * i) for i in range(len(stack_words))]

    def _exec(self, ins):
        mn, ops = ins.mn, ins.ops
        nxt = ins.next
        if mn == "movq" and (ops[0][0] == "x" or ops[1][0] == "x"):
            if ops[0][0] == "x":
                self.put(ops[1], 8, self.x[ops[0][1]] & M64)
            else:
                self.x[ops[1][1]] = self.get(ops[0], 8)
            return nxt
        if mn == "mov" or mn in ("movq", "movl", "movb", "movw", "movabs"):
            size = self.opsize(ops, mn)
            self.put(ops[1], size, self.get(ops[0], size))
            return nxt
        if len(mn) == 6 and mn[:4] in ("movz", "movs") and mn[4] in "bwl" and mn[5] in "wlq":
            ss = {"b": 1, "w": 2, "l": 4}[mn[4]]
            v = self.get(ops[0], ss)
            if mn[3] == "s":
                v = _sx(v, ss)
            self.put(ops[1], {"w": 2, "l": 4, "q": 8}[mn[5]], v)
            return nxt
        if mn.startswith("set") and len(ops) == 1:
            self.put(ops[0], 1, int(self.cond(mn[3:])))
            return nxt
        if mn == "lea":
            a = ops[0][1] if ops[0][0] == "c" else self.ea(ops[0])
            self.put(ops[1], ops[1][2], a)
            return nxt
        if mn in ("add", "addq", "addl", "sub", "subq", "subl", "cmp", "cmpq", "cmpl", "cmpb", "adc", "sbb"):
            size = self.opsize(ops, mn)
            src, dst = self.get(ops[0], size), self.get(ops[1], size)
            if ops[0][0] == "i":
                src = _sx(src, size) & ((1 << (size * 8)) - 1) if size == 8 else src
            base = mn.rstrip("qlbw") if mn not in ("sbb", "sub") else mn
            if base.startswith("add"):
                self.put(ops[1], size, self.setflags_add(dst, src, size))
            elif base == "adc":
                self.put(ops[1], size, self.setflags_add(dst, src, size, self.cf))
            elif base.startswith("sub"):
                self.put(ops[1], size, self.setflags_sub(dst, src, size))
            elif base == "sbb":
                self.put(ops[1], size, self.setflags_sub(dst, src, size, self.cf))
            else:
                self.setflags_sub(dst, src, size)
            return nxt
        if mn == "mul":
            size = self.opsize(ops, mn)
            assert size == 8
            p = self.r[0] * self.get(ops[0], 8)
            self.r[0], self.r[2] = p & M64, p >> 64
            self.cf = self.of = int(self.r[2] != 0)
            return nxt
        if mn == "div":
            size = self.opsize(ops, mn)
            assert size == 8
            d = self.get(ops[0], 8)
            n = (self.r[2] << 64) | self.r[0]
            if d == 0 or n // d > M64:
                raise ZeroDivisionError("div fault at %x" % ins.addr)
            self.r[0], self.r[2] = n // d, n % d
            return nxt
        if mn == "cqto":
            self.r[2] = M64 if self.r[0] >> 63 else 0
            return nxt
        if mn == "idiv":
            size = self.opsize(ops, mn)
            assert size == 8
            d = _sx(self.get(ops[0], 8), 8)
            n = _sx((self.r[2] << 64) | self.r[0], 16)
            if d == 0:
                raise ZeroDivisionError("idiv fault at %x" % ins.addr)
            qt = abs(n) // abs(d)
            qt = -qt if (n < 0) != (d < 0) else qt
            self.r[0], self.r[2] = qt & M64, (n - qt * d) & M64
            return nxt
        if mn == "bswap":
            size = self.opsize(ops, mn)
            v = self.get(ops[0], size)
            self.put(ops[0], size, int.from_bytes(v.to_bytes(size, "little"), "big"))
            return nxt
        if mn in ("rol", "ror"):
            size = self.opsize(ops[-1:], mn)
            cnt, dst = (1, ops[0]) if len(ops) == 1 else (self.get(ops[0], 1) & (size * 8 - 1), ops[1])
            v, bits = self.get(dst, size), size * 8
            if cnt:
                res = ((v << cnt) | (v >> (bits - cnt))) if mn == "rol" else ((v >> cnt) | (v << (bits - cnt)))
                self.put(dst, size, res & ((1 << bits) - 1))
            return nxt
        if mn == "neg":
            size = self.opsize(ops, mn)
            v = self.get(ops[0], size)
            res = self.setflags_sub(0, v, size)
            self.put(ops[0], size, res)
            return nxt
        if mn == "not":
            size = self.opsize(ops, mn)
            self.put(ops[0], size, ~self.get(ops[0], size))
            return nxt
        if mn == "imul":
            size = self.opsize(ops, mn)
            assert len(ops) == 2
            res = (self.get(ops[0], size) * self.get(ops[1], size)) & ((1 << (size * 8)) - 1)
            self.put(ops[1], size, res)
            return nxt
        if mn.startswith("cmov"):
            size = self.opsize(ops, mn)
            if self.cond(mn[4:]):
                self.put(ops[1], size, self.get(ops[0], size))
            elif size == 4 and ops[1][0] == "r":
                self.r[ops[1][1]] &= 0xffffffff
            return nxt
        if mn == "jmp":
            return ins.target if ins.target is not None else self.get(ops[0], 8)
        if mn == "call":
            tgt = ins.target if ins.target is not None else self.get(ops[0], 8)
            if self.tracers and tgt in self.tracers:
                self.tracers[tgt](self)   # observer only: sees the argument area at [rsp], then the call proceeds
            if tgt in self.hooks:
                self.hooks[tgt](self)
                return nxt
            if tgt not in self.code and not self.autoload(tgt):
                raise RuntimeError("call to non-whitelisted %x (%s) at %x" % (tgt, self.name_of(tgt), ins.addr))
            self.r[4] = (self.r[4] - 8) & M64
            self.wq(self.r[4], nxt)
            return tgt
        if mn == "ret":
            pc = self.rq(self.r[4])
            self.r[4] = (self.r[4] + 8) & M64
            return pc
        if mn[0] == "j":
            return ins.target if self.cond(mn[1:]) else nxt
        if mn.startswith("nop") or mn == "xchg" and ops[0] == ops[1]:
            return nxt
        if mn == "xchg":
            size = self.opsize(ops, mn)
            a, b = self.get(ops[0], size), self.get(ops[1], size)
            self.put(ops[0], size, b)
            self.put(ops[1], size, a)
            return nxt
        if mn in ("xor", "or", "and", "test", "testb", "testq", "andq", "orq"):
            size = self.opsize(ops, mn)
            a, b = self.get(ops[0], size), self.get(ops[1], size)
            if ops[0][0] == "i":
                a = _sx(a if size < 8 else a, size) & ((1 << (size * 8)) - 1)
            base = mn.rstrip("qb") if mn not in ("xor", "or", "and", "test") else mn
            res = a ^ b if base == "xor" else (a | b if base == "or" else a & b)
            self.setflags_logic(res, size)
            if base != "test":
                self.put(ops[1], size, res)
            return nxt
        if mn in ("shl", "shr", "sar"):
            size = self.opsize(ops[-1:], mn)
            if len(ops) == 1:
                cnt, dst = 1, ops[0]
            else:
                cnt, dst = self.get(ops[0], 1) & (63 if size == 8 else 31), ops[1]
            v = self.get(dst, size)
            bits = size * 8
            if cnt:
                if mn == "shl":
                    self.cf = (v >> (bits - cnt)) & 1
                    res = (v << cnt) & ((1 << bits) - 1)
                elif mn == "shr":
                    self.cf = (v >> (cnt - 1)) & 1
                    res = v >> cnt
                else:
                    self.cf = (v >> (cnt - 1)) & 1
                    res = (_sx(v, size) >> cnt) & ((1 << bits) - 1)
                self.zf = int(res == 0)
                self.sf = res >> (bits - 1)
                self.put(dst, size, res)
            return nxt
        if mn == "inc" or mn == "dec":
            size = self.opsize(ops, mn)
            cf = self.cf
            v = self.get(ops[0], size)
            res = self.setflags_add(v, 1, size) if mn == "inc" else self.setflags_sub(v, 1, size)
            self.cf = cf
            self.put(ops[0], size, res)
            return nxt
        if mn in ("bt", "bts"):
            size = self.opsize(ops, mn)
            bit = self.get(ops[0], 1) & (size * 8 - 1)
            v = self.get(ops[1], size)
            self.cf = (v >> bit) & 1
            if mn == "bts":
                self.put(ops[1], size, v | (1 << bit))
            return nxt
        if mn == "bsr":
            size = self.opsize(ops, mn)
            v = self.get(ops[0], size)
            self.zf = int(v == 0)
            if v:
                self.put(ops[1], size, v.bit_length() - 1)
            return nxt
        # ---- SSE scalar double ----
        if mn == "movsd":
            if ops[0][0] == "x" and ops[1][0] == "x":
                self.x[ops[1][1]] = (self.x[ops[1][1]] & ~M64) | (self.x[ops[0][1]] & M64)
            elif ops[0][0] == "x":
                self.wq(self.ea(ops[1]), self.x[ops[0][1]] & M64)
            else:
                self.x[ops[1][1]] = self.get(ops[0], 8)
            return nxt
        if mn in ("movups", "movaps", "movupd", "movapd", "movdqu", "movdqa"):
            if ops[0][0] == "x" and ops[1][0] == "x":
                self.x[ops[1][1]] = self.x[ops[0][1]]
            elif ops[0][0] == "x":
                a = self.ea(ops[1])
                self.wq(a, self.x[ops[0][1]] & M64)
                self.wq(a + 8, self.x[ops[0][1]] >> 64)
            else:
                a = self.ea(ops[0])
                self.x[ops[1][1]] = self.rq(a) | (self.rq(a + 8) << 64)
            return nxt
        if mn in ("xorps", "xorpd", "pxor", "andpd", "andps", "pand", "orpd", "orps", "por", "andnpd", "andnps", "pandn"):
            a = self.x[ops[0][1]] if ops[0][0] == "x" else (self.get(ops[0], 8) | (self.rq((ops[0][1] if ops[0][0] == "c" else self.ea(ops[0])) + 8) << 64))
            d = self.x[ops[1][1]]
            if mn in ("xorps", "xorpd", "pxor"):
                d ^= a
            elif mn in ("andpd", "andps", "pand"):
                d &= a
            elif mn in ("orpd", "orps", "por"):
                d |= a
            else:
                d = (~d & ((1 << 128) - 1)) & a
            self.x[ops[1][1]] = d
            return nxt
        if mn in ("cmpltsd", "cmplesd", "cmpeqsd", "cmpneqsd", "cmpnltsd", "cmpnlesd", "cmpunordsd", "cmpordsd"):
            b = b2f(self.x[ops[0][1]] if ops[0][0] == "x" else self.get(ops[0], 8))
            a = b2f(self.x[ops[1][1]])
            un = a != a or b != b
            r = {"cmpltsd": (not un) and a < b, "cmplesd": (not un) and a <= b, "cmpeqsd": (not un) and a == b,
                 "cmpneqsd": un or a != b, "cmpnltsd": un or not a < b, "cmpnlesd": un or not a <= b,
                 "cmpunordsd": un, "cmpordsd": not un}[mn]
            self.x[ops[1][1]] = (self.x[ops[1][1]] & ~M64) | (M64 if r else 0)
            return nxt
        if mn == "sqrtsd":
            b = b2f(self.x[ops[0][1]] if ops[0][0] == "x" else self.get(ops[0], 8))
            import math
            self.x[ops[1][1]] = (self.x[ops[1][1]] & ~M64) | f2b(math.sqrt(b) if b >= 0 else float("nan"))
            return nxt
        if mn == "btr":
            size = self.opsize(ops[-1:], mn)
            bit = self.get(ops[0], 1) & (size * 8 - 1)
            v = self.get(ops[1], size)
            self.cf = (v >> bit) & 1
            self.put(ops[1], size, v & ~(1 << bit))
            return nxt
        if mn in ("addsd", "subsd", "divsd", "mulsd"):
            b = b2f(self.x[ops[0][1]] if ops[0][0] == "x" else self.get(ops[0], 8))
            a = b2f(self.x[ops[1][1]])
            res = a + b if mn == "addsd" else (a - b if mn == "subsd" else (a / b if mn == "divsd" else a * b))
            self.x[ops[1][1]] = (self.x[ops[1][1]] & ~M64) | f2b(res)
            return nxt
        if mn in ("cvtsi2sd", "cvtsi2sdq", "cvtsi2sdl"):
            sz = ops[0][2] if ops[0][0] == "r" else (4 if mn.endswith("l") else 8)
            v = _sx(self.get(ops[0], sz), sz)
            self.x[ops[1][1]] = (self.x[ops[1][1]] & ~M64) | f2b(float(v))
            return nxt
        if mn == "cvttsd2si":
            f = b2f(self.x[ops[0][1]])
            v = int(f) if abs(f) < 9.3e18 else -(1 << 63)
            self.put(ops[1], 8, v & M64)
            return nxt
        if mn == "ucomisd":
            b = b2f(self.x[ops[0][1]] if ops[0][0] == "x" else self.get(ops[0], 8))
            a = b2f(self.x[ops[1][1]])
            if a != a or b != b:
                self.zf = self.pf = self.cf = 1
            else:
                self.zf, self.pf, self.cf = int(a == b), 0, int(a < b)
            self.of = self.sf = 0
            return nxt
        raise NotImplementedError("%x: %s %s" % (ins.addr, mn, ops))
