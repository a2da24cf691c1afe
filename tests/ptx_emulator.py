"""A small interpreter for the PTX subset used by the generated clause-loop handlers
(tools/gen_interval_loop.py): enough to run a handler's arithmetic on the CPU, one lane at a time,
with IEEE directed rounding done in exact rational arithmetic.  Test infrastructure only.

Registers hold raw 32-bit patterns; float instructions reinterpret them.  Supported: mov, neg, abs,
add/sub/mul/div/sqrt with .rm / .rp / .rn, fma.rn, mul.ftz, rcp / rsqrt.approx (modelled as the correctly
rounded value; RSQ_ULPS moves the latter), min, max, setp (f32 and u32, with a combining predicate), the
warp vote of one lane, selp, and/or/not on predicates, and/or/shl/shr/add on b32, mad.wide.u32, predicated st.u32 / mov (the verdict record)."""
from __future__ import annotations

import math
import re
import struct
from fractions import Fraction

import numpy as np

FLT_MAX = float(np.finfo(np.float32).max)


def f2b(x) -> int:
    return struct.unpack("<I", struct.pack("<f", np.float32(x)))[0]


def b2f(b: int) -> np.float32:
    return np.frombuffer(struct.pack("<I", b & 0xffffffff), dtype="<f4")[0]


def _round(fr: Fraction, mode: str) -> np.float32:
    """Exact rational -> float32 under rm (toward -inf), rp (toward +inf) or rn."""
    if fr == 0:
        return np.float32(0.0)
    with np.errstate(over="ignore", under="ignore"):
        c = np.float32(float(fr)) if abs(fr) < Fraction(10) ** 60 else np.float32(math.copysign(math.inf, fr))
    if mode == "rn":
        # float(fr) is correctly rounded to double; double -> float32 can double-round only on exact
        # ties of the float32 grid, which the callers' operands (float32 sums / products) cannot produce
        # except through the paths below - resolve by exact comparison of the two neighbours.
        if math.isinf(c):
            return c
        lo = c if Fraction(float(c)) <= fr else np.nextafter(c, np.float32(-np.inf))
        hi = np.nextafter(lo, np.float32(np.inf))
        if math.isinf(hi):
            return lo if fr - Fraction(float(lo)) < Fraction(float(np.finfo(np.float32).max)) * Fraction(1, 2 ** 25) else hi
        dl, dh = fr - Fraction(float(lo)), Fraction(float(hi)) - fr
        if dl != dh:
            return lo if dl < dh else hi
        return lo if (f2b(lo) & 1) == 0 else hi
    inf, ninf = np.float32(np.inf), np.float32(-np.inf)
    if math.isinf(c):
        c = np.float32(FLT_MAX if c > 0 else -FLT_MAX)
    with np.errstate(over="ignore"):
        lo = c
        while not math.isinf(lo) and Fraction(float(lo)) > fr:        # step down to a float32 <= fr
            lo = np.nextafter(lo, ninf)
        if math.isinf(lo):                                            # fr < -FLT_MAX
            return ninf if mode == "rm" else np.float32(-FLT_MAX)
        while True:                                                   # climb to the last float32 <= fr
            n = np.nextafter(lo, inf)
            if math.isinf(n) or Fraction(float(n)) > fr:
                break
            lo = n
        if mode == "rm" or Fraction(float(lo)) == fr:
            return lo
        return np.nextafter(lo, inf)                                  # +inf beyond FLT_MAX


def _zero(sign_negative: bool) -> np.float32:
    return np.float32(-0.0) if sign_negative else np.float32(0.0)


def fadd(a, b, mode):
    a, b = np.float32(a), np.float32(b)
    if np.isnan(a) or np.isnan(b):
        return np.float32(np.nan)
    if np.isinf(a) or np.isinf(b):
        if np.isinf(a) and np.isinf(b) and a != b:
            return np.float32(np.nan)
        return a if np.isinf(a) else b
    s = Fraction(float(a)) + Fraction(float(b))
    if s == 0:
        if a == 0 and b == 0 and np.signbit(a) == np.signbit(b):
            return a
        return _zero(mode == "rm")
    return _round(s, mode)


def fmul(a, b, mode):
    a, b = np.float32(a), np.float32(b)
    if np.isnan(a) or np.isnan(b):
        return np.float32(np.nan)
    neg = bool(np.signbit(a)) != bool(np.signbit(b))
    if np.isinf(a) or np.isinf(b):
        if a == 0 or b == 0:
            return np.float32(np.nan)
        return np.float32(-np.inf if neg else np.inf)
    if a == 0 or b == 0:
        return _zero(neg)
    r = _round(Fraction(float(a)) * Fraction(float(b)), mode)
    return _zero(neg) if r == 0 else r


def fdiv(a, b, mode):
    a, b = np.float32(a), np.float32(b)
    if np.isnan(a) or np.isnan(b):
        return np.float32(np.nan)
    neg = bool(np.signbit(a)) != bool(np.signbit(b))
    if np.isinf(a):
        return np.float32(np.nan) if np.isinf(b) else np.float32(-np.inf if neg else np.inf)
    if np.isinf(b):
        return _zero(neg)
    if b == 0:
        return np.float32(np.nan) if a == 0 else np.float32(-np.inf if neg else np.inf)
    if a == 0:
        return _zero(neg)
    r = _round(Fraction(float(a)) / Fraction(float(b)), mode)
    return _zero(neg) if r == 0 else r


def fsqrt(a, mode):
    a = np.float32(a)
    if np.isnan(a) or a < 0:
        return np.float32(np.nan)
    if a == 0 or np.isinf(a):
        return a
    x = Fraction(float(a))
    c = np.float32(math.sqrt(float(a)))
    while Fraction(float(c)) ** 2 > x:
        c = np.nextafter(c, np.float32(-np.inf))
    while Fraction(float(np.nextafter(c, np.float32(np.inf)))) ** 2 <= x:
        c = np.nextafter(c, np.float32(np.inf))
    if Fraction(float(c)) ** 2 == x or mode == "rm":
        return c
    if mode == "rp":
        return np.nextafter(c, np.float32(np.inf))
    n = np.nextafter(c, np.float32(np.inf))
    mid = (Fraction(float(c)) + Fraction(float(n))) / 2
    return c if mid * mid > x else n


def ffma(a, b, c):
    """fma.rn.f32 on finite operands (the hand-expanded sqrt / division sequences of the float loop):
    one rounding of the exact a b + c."""
    a, b, c = np.float32(a), np.float32(b), np.float32(c)
    if not (np.isfinite(a) and np.isfinite(b) and np.isfinite(c)):
        with np.errstate(all="ignore"):
            return np.float32(np.float64(a) * np.float64(b) + np.float64(c))   # NaN / infinity propagation only
    s = Fraction(float(a)) * Fraction(float(b)) + Fraction(float(c))
    if s == 0:
        pneg = bool(np.signbit(a)) != bool(np.signbit(b))
        return _zero(pneg and bool(np.signbit(c)))
    return _round(s, "rn")


# MUFU.RSQ / MUFU.RCP are approximations (about one unit in the last place): the model returns the correctly
# rounded value moved by RSQ_ULPS / RCP_ULPS units.  The sqrt sequence of the float loops gives the correctly
# rounded result for any value that close (tests run -2 .. 2); the division sequence for a reciprocal within one
# unit, EXCEPT for divisors whose mantissa is all ones (there one Newton step leaves the reciprocal a unit off and
# the quotient with it - measured with this model), which is why the float loops send those to div.rn.
RSQ_ULPS = 0
RCP_ULPS = 0


def _nudge(x: np.float32, ulps: int) -> np.float32:
    for _ in range(abs(ulps)):
        x = np.nextafter(x, np.float32(np.inf if ulps > 0 else -np.inf))
    return x


def _ftz(x: np.float32) -> np.float32:
    x = np.float32(x)
    if x != 0 and np.isfinite(x) and abs(x) < np.finfo(np.float32).tiny:
        return _zero(bool(np.signbit(x)))
    return x


def frcp_approx(a):
    a = _ftz(a)
    if np.isnan(a) or a == 0 or np.isinf(a):
        with np.errstate(all="ignore"):
            return np.float32(1.0) / a
    return _ftz(_nudge(fdiv(np.float32(1.0), a, "rn"), RCP_ULPS))


def frsqrt_approx(a):
    a = _ftz(a)
    if np.isnan(a) or a < 0:
        return np.float32(np.nan)
    if a == 0:
        return np.float32(np.inf)
    if np.isinf(a):
        return np.float32(0.0)
    return _ftz(_nudge(np.float32(1.0 / math.sqrt(float(a))), RSQ_ULPS))     # double precision, then one rounding: within half a unit


def fmin(a, b):
    a, b = np.float32(a), np.float32(b)
    if np.isnan(a):
        return b
    if np.isnan(b):
        return a
    return a if a < b or (a == b and np.signbit(a)) else b


def fmax(a, b):
    a, b = np.float32(a), np.float32(b)
    if np.isnan(a):
        return b
    if np.isnan(b):
        return a
    return a if a > b or (a == b and not np.signbit(a)) else b


class Machine:
    def __init__(self, regs=None, operands=None):
        self.r = dict(regs or {})       # name -> uint32 pattern (or bool for predicates, int for b64)
        self.ops = dict(operands or {})  # "%3" -> name of the register that stands for it
        self.stores = []                # (address, value) of st.u32

    def val(self, tok):
        tok = tok.strip()
        if tok in self.ops:
            tok = self.ops[tok]
        if tok.startswith("0f"):
            return int(tok[2:], 16)
        if re.fullmatch(r"-?\d+", tok):
            return int(tok) & 0xffffffff
        if re.fullmatch(r"0x[0-9a-fA-F]+", tok):
            return int(tok, 16)
        return self.r[tok]

    def set(self, tok, v):
        tok = tok.strip()
        self.r[self.ops.get(tok, tok)] = v

    def run(self, text: str):
        for ins in [i.strip() for i in text.split(";") if i.strip()]:
            pred = None
            m = re.match(r"@(!?)(\w+)\s+(.*)", ins)
            if m:
                pred = bool(self.r[m.group(2)]) != bool(m.group(1))
                ins = m.group(3)
                if not pred:
                    continue
            op, rest = ins.split(None, 1)
            a = [x.strip() for x in rest.replace("[", "").replace("]", "").split(",")]
            f = lambda k: b2f(self.val(a[k]))
            if op in ("mov.b32", "mov.f32", "mov.u32"):
                self.set(a[0], self.val(a[1]))
            elif op == "neg.f32":
                self.set(a[0], self.val(a[1]) ^ 0x80000000)
            elif op == "abs.f32":
                self.set(a[0], self.val(a[1]) & 0x7fffffff)
            elif re.fullmatch(r"(add|sub|mul|div)\.(rm|rp|rn)\.f32", op):
                kind, mode, _ = op.split(".")
                x, y = f(1), f(2)
                if kind == "sub":
                    y = np.float32(-y) if not np.isnan(y) else y
                fn = {"add": fadd, "sub": fadd, "mul": fmul, "div": fdiv}[kind]
                self.set(a[0], f2b(fn(x, y, mode)))
            elif re.fullmatch(r"sqrt\.(rm|rp|rn)\.f32", op):
                self.set(a[0], f2b(fsqrt(f(1), op.split(".")[1])))
            elif op == "min.f32":
                self.set(a[0], f2b(fmin(f(1), f(2))))
            elif op == "max.f32":
                self.set(a[0], f2b(fmax(f(1), f(2))))
            elif re.fullmatch(r"setp\.(lt|gt|le|ge|eq|ne)\.f32", op):
                x, y = f(1), f(2)
                c = op.split(".")[1]
                self.set(a[0], bool({"lt": x < y, "gt": x > y, "le": x <= y, "ge": x >= y, "eq": x == y, "ne": x != y}[c]))
            elif re.fullmatch(r"setp\.(lt|gt|le|ge|eq|ne)\.(or|and)\.u32", op):
                x, y = self.val(a[1]), self.val(a[2])
                c, bop = op.split(".")[1:3]
                t = {"lt": x < y, "gt": x > y, "le": x <= y, "ge": x >= y, "eq": x == y, "ne": x != y}[c]
                self.set(a[0], (t or bool(self.r[a[3]])) if bop == "or" else (t and bool(self.r[a[3]])))
            elif op == "vote.sync.any.pred":
                self.set(a[0], bool(self.r[a[1]]))          # one lane stands for the warp
            elif op == "fma.rn.f32":
                self.set(a[0], f2b(ffma(f(1), f(2), f(3))))
            elif op == "mul.ftz.f32":
                self.set(a[0], f2b(_ftz(fmul(_ftz(f(1)), _ftz(f(2)), "rn"))))
            elif op == "rcp.approx.ftz.f32":
                self.set(a[0], f2b(frcp_approx(f(1))))
            elif op == "rsqrt.approx.ftz.f32":
                self.set(a[0], f2b(frsqrt_approx(f(1))))
            elif re.fullmatch(r"setp\.(lt|gt|le|ge|eq|ne)\.u32", op):
                x, y = self.val(a[1]), self.val(a[2])
                c = op.split(".")[1]
                self.set(a[0], {"lt": x < y, "gt": x > y, "le": x <= y, "ge": x >= y, "eq": x == y, "ne": x != y}[c])
            elif op == "selp.b32":
                self.set(a[0], self.val(a[1]) if self.r[a[3]] else self.val(a[2]))
            elif op == "and.pred":
                self.set(a[0], bool(self.r[a[1]]) and bool(self.r[a[2]]))
            elif op == "or.pred":
                self.set(a[0], bool(self.r[a[1]]) or bool(self.r[a[2]]))
            elif op == "not.pred":
                self.set(a[0], not bool(self.r[a[1]]))
            elif op == "and.b32":
                self.set(a[0], self.val(a[1]) & self.val(a[2]))
            elif op == "or.b32":
                self.set(a[0], self.val(a[1]) | self.val(a[2]))
            elif op == "shl.b32":
                s = self.val(a[2])
                self.set(a[0], (self.val(a[1]) << s) & 0xffffffff if s < 32 else 0)
            elif op == "shr.u32":
                s = self.val(a[2])
                self.set(a[0], self.val(a[1]) >> s if s < 32 else 0)
            elif op == "add.u32":
                self.set(a[0], (self.val(a[1]) + self.val(a[2])) & 0xffffffff)
            elif op == "mad.wide.u32":
                self.set(a[0], self.val(a[1]) * self.val(a[2]) + self.val(a[3]))
            elif op == "st.u32":
                self.stores.append((self.val(a[0]), self.val(a[1])))
            else:
                raise NotImplementedError(ins)
        return self


# ---- whole-loop execution: labels, the branch table, shared memory ---------------------------------

def load_asm(path) -> str:
    """The asm template of a generated *_ptx.inc as one string, local-label suffixes removed."""
    text = "".join(re.findall(r'"((?:[^"\\]|\\.)*)"', open(path).read()))
    return text.replace("\\n", "\n").replace("_%=", "")


class LoopMachine(Machine):
    """Runs a complete generated clause loop for ONE lane: `smem` maps 4-byte-aligned shared
    addresses to 32-bit words; operands %0.. are bound to registers by name."""

    def __init__(self, asm: str, regs, operands, smem):
        super().__init__(regs, operands)
        self.smem = smem
        self.tmem = {}                  # tensor memory of this lane: column -> 32-bit word
        body = asm[asm.index("{") + 1: asm.rindex("}")]
        m = re.search(r"(\w+):\s*\.branchtargets([^;]*);", body)
        self.table = [t.strip() for t in m.group(2).replace("\n", " ").split(",")]
        body = body[:m.start()] + body[m.end():]
        self.prog, self.labels = [], {}
        for stmt in body.split(";"):
            stmt = stmt.strip()
            while True:
                lm = re.match(r"(\w+):\s*(.*)", stmt, re.S)
                if not lm:
                    break
                self.labels[lm.group(1)] = len(self.prog)
                stmt = lm.group(2).strip()
            if stmt and not stmt.startswith(".reg"):
                self.prog.append(stmt)

    def execute(self, max_steps=200000):
        pc = 0
        for _ in range(max_steps):
            if pc >= len(self.prog):
                return self
            ins = self.prog[pc]
            pc += 1
            pm = re.match(r"@(!?)(\w+)\s+(.*)", ins, re.S)
            if pm:
                if bool(self.r[pm.group(2)]) == bool(pm.group(1)):
                    continue
                ins = pm.group(3)
            op = ins.split(None, 1)[0]
            if op == "bra.uni":
                pc = self.labels[ins.split()[1]]
            elif op == "brx.idx.uni":
                idx = self.val(ins.split(None, 1)[1].split(",")[0])
                pc = self.labels[self.table[idx]]
            elif op == "ld.shared.v2.b32":
                m = re.match(r"ld\.shared\.v2\.b32\s*\{(.+?),(.+?)\}\s*,\s*\[(.+?)\]", ins)
                addr = self.val(m.group(3))
                self.set(m.group(1), self.smem.get(addr, 0))
                self.set(m.group(2), self.smem.get(addr + 4, 0))
            elif op == "st.shared.v2.b32":
                m = re.match(r"st\.shared\.v2\.b32\s*\[(.+?)\]\s*,\s*\{(.+?),(.+?)\}", ins)
                addr = self.val(m.group(1))
                self.smem[addr] = self.val(m.group(2))
                self.smem[addr + 4] = self.val(m.group(3))
            elif op.startswith("tcgen05.wait::"):
                pass                                          # completion of asynchronous tensor-memory accesses
            elif op.startswith("tcgen05.ld.") or op.startswith("tcgen05.st."):
                # tensor memory as this lane sees it: 32-bit columns addressed by the column field
                am = re.search(r"\[(\w+|%\d+)\]", ins)
                col = self.val(am.group(1))
                names = [x.strip() for x in re.search(r"\{(.*?)\}", ins).group(1).split(",")]
                for k, name in enumerate(names):
                    if op.startswith("tcgen05.ld."):
                        self.set(name, self.tmem.get(col + k, 0))
                    else:
                        self.tmem[col + k] = self.val(name)
            elif op == "bfe.u32":
                a = [x.strip() for x in ins.split(None, 1)[1].split(",")]
                self.set(a[0], (self.val(a[1]) >> self.val(a[2])) & ((1 << self.val(a[3])) - 1))
            elif op == "ld.shared.b32":
                am = re.search(r"\[(\w+|%\d+)(?:\+(\d+))?\]", ins)
                addr = self.val(am.group(1)) + int(am.group(2) or 0)
                self.set(ins.split(None, 1)[1].split(",")[0], self.smem.get(addr, 0))
            elif op in ("ld.shared.b64", "ld.shared.v2.b64", "st.shared.b64", "st.shared.v2.b64"):
                # 64-bit registers hold two f32 patterns (lo | hi << 32): the float loop's sample pairs
                am = re.search(r"\[(\w+|%\d+)(?:\+(\d+))?\]", ins)
                addr = self.val(am.group(1)) + int(am.group(2) or 0)
                names = [x.strip() for x in re.sub(r"\[.*?\]", "", ins.split(None, 1)[1]).replace("{", "").replace("}", "").split(",") if x.strip()]
                for k, name in enumerate(names):
                    if op.startswith("ld"):
                        self.set(name, self.smem.get(addr + 8 * k, 0) | (self.smem.get(addr + 8 * k + 4, 0) << 32))
                    else:
                        v = self.r[name]
                        self.smem[addr + 8 * k] = v & 0xffffffff
                        self.smem[addr + 8 * k + 4] = v >> 32
            elif op == "mov.b64":
                body = ins.split(None, 1)[1]
                um = re.fullmatch(r"\{(\w+),\s*(\w+)\}\s*,\s*(\w+)", body.strip())
                pm2 = re.fullmatch(r"(\w+)\s*,\s*\{(\w+),\s*(\w+)\}", body.strip())
                if um:                                       # unpack
                    v = self.r[um.group(3)]
                    self.set(um.group(1), v & 0xffffffff)
                    self.set(um.group(2), v >> 32)
                elif pm2:                                    # pack
                    self.set(pm2.group(1), self.val(pm2.group(2)) | (self.val(pm2.group(3)) << 32))
                else:
                    d, src = [x.strip() for x in body.split(",")]
                    self.set(d, 0 if src == "0" else self.r[src])
            elif re.fullmatch(r"(add|sub|mul)\.rn\.f32x2", op):
                d, x, y = [t.strip() for t in ins.split(None, 1)[1].split(",")]
                kind = op.split(".")[0]
                out = 0
                for half in (0, 32):
                    xa, yb = b2f((self.r[x] >> half) & 0xffffffff), b2f((self.r[y] >> half) & 0xffffffff)
                    if kind == "sub":
                        yb = np.float32(-yb) if not np.isnan(yb) else yb
                    out |= f2b({"add": fadd, "sub": fadd, "mul": fmul}[kind](xa, yb, "rn")) << half
                self.set(d, out)
            elif op == "prmt.b32":
                a = [x.strip() for x in ins.split(None, 1)[1].split(",")]
                src = (self.val(a[2]) << 32) | self.val(a[1])
                sel, out = self.val(a[3]), 0
                for k in range(4):
                    out |= ((src >> (8 * ((sel >> (4 * k)) & 7))) & 0xff) << (8 * k)
                self.set(a[0], out)
            else:
                self.run(ins)
        raise RuntimeError("loop did not finish")
