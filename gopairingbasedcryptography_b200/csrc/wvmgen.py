"""Warp-VM program generator: ONE WARP (32 lanes) per pairing, Fp-level micro-ops -- the latency path.

Why: the reference calls bn254.Pair with 1-element slices (60 call sites) and its BASELINE configs 0 / 4 are small
batches (1024 BLS messages; 64 AFP25 ciphertexts x 3 pairs).  One thread needs ~20 ms per pairing and the K = 3
lane-group kernel ~6 ms: both are latency-bound by a ~15 000-long chain of dependent Fp products.  A pairing has far
more parallelism than that: an Fp12 product is 54 independent Fp products.  Here the pairing formulas (the SAME traced
DAG as the lane-group VM, vmgen.py) are lowered to Fp level and scheduled 32 wide:

  * every Fp value is a LINEAR FORM  sum c_i * base_i  over "bases" (inputs, constants, Fp products, one inversion)
    with small integer coefficients; additions, subtractions, negations, conjugations, xi-multiplications and the
    Karatsuba recombinations never become instructions -- they only edit the form;
  * a form is MATERIALISED (one LIN op: up to 15 terms, one reduction) only where a product needs it as an operand,
    where it grows past the term cap, or at the outputs;
  * MUL computes (s0 +- s1) * (s2 +- s3) on canonical slots (the Karatsuba pre-additions ride in the operands);
  * rounds hold <= 32 ops of ONE class (MUL / LIN / INV), list-scheduled by critical path; one __syncwarp() per round.

Emits wvm_prog_<name>.inc (16 x u16 per op) + wvm_prog_meta.cuh, and can evaluate a program on Python integers.
Self-contained apart from vmgen.py's tracer (no oracle import).  Run:  python wvmgen.py
"""
from __future__ import annotations

import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import vmgen as g  # noqa: E402

P = g.P
LANES = 32
TMAX = int(os.environ.get("WVM_TMAX", "8"))       # terms per LIN op (encoding limit: 15); 8 measured best by the cost model
CMAX = 31                                         # |coefficient| limit (6-bit signed field)
CSUM_MAX = 120                                    # sum |c| per LIN: 128 p + sum c_i s_i stays inside (0, 256 p)
SHARE = os.environ.get("WVM_SHARE", "1") == "1"   # materialise values with several consumers instead of inlining their forms
POLICY = os.environ.get("WVM_POLICY", "fill12")
DUP_MAX = int(os.environ.get("WVM_DUP_MAX", "6"))   # a LIN of at most this many terms may be duplicated into several consumers
# term cap when LIN -> LIN chains are collapsed (0 = off).  Measured on B200 (benchmarks/wvm_ab.py, profiles/r2/wvm_ab.jsonl):
# collapsing helps the Miller program (0.705 -> 0.672 ms) and hurts the final exponentiation (0.76 -> 0.86 ms).
TINLINE_ENV = os.environ.get("WVM_TINLINE")
TINLINE_BY_PROGRAM = {"miller": 15, "miller2": 15, "finalexp": 0, "pair": 0}
TINLINE = 0
SHARE_MIN = int(os.environ.get("WVM_SHARE_MIN", "3"))  # ... when the form has at least this many terms
OP_NOP, OP_MUL, OP_LIN, OP_INV = 0, 1, 2, 3
COST = {OP_MUL: 1.0, OP_LIN: 0.35, OP_INV: 40.0}
INV2 = pow(2, -1, P)


SPLIT_BIG = os.environ.get("WVM_SPLIT_BIG", "0") == "1"  # measured: no fewer LIN rounds (profiles/r2/wvm_generator_shapes.txt)


def nfields(f):
    """Term fields a form occupies in one LIN record: a coefficient beyond the 6-bit field is written as several terms on
    the same slot (54 b = 31 b + 23 b) instead of materialising 31 b in a LIN of its own (one dependent round less)."""
    vals = f.values() if isinstance(f, dict) else [c for _, c in f]
    if not SPLIT_BIG:
        return len(vals) if all(abs(c) <= CMAX for c in vals) else 99
    return sum(max(1, -(-abs(c) // CMAX)) for c in vals)


class Base:
    """A materialised Fp value living in a slot."""
    __slots__ = ("id", "kind", "op", "src", "flags", "terms", "value_const", "users", "round", "slot", "prio", "name", "depth")

    def __init__(self, id, kind):
        self.id, self.kind = id, kind  # kind: "in", "const", "op"
        self.op, self.src, self.flags, self.terms = OP_NOP, (), 0, ()
        self.value_const, self.users, self.round, self.slot, self.prio, self.name, self.depth = None, [], None, None, 0.0, None, 0


class Lowering:
    def __init__(self):
        self.bases = []
        self.consts = {}
        self.lin_cache = {}
        self.mul_cache = {}
        self.zero = self.const(0)

    def _new(self, kind):
        b = Base(len(self.bases), kind)
        self.bases.append(b)
        return b

    def input(self, name):
        b = self._new("in")
        b.name = name
        return b

    def const(self, v):
        v %= P
        if v not in self.consts:
            b = self._new("const")
            b.value_const = v
            self.consts[v] = b
        return self.consts[v]

    # ---- forms: dict base_id -> coef ------------------------------------------------------------------------
    @staticmethod
    def f_add(x, y, cy=1):
        out = dict(x)
        for k, c in y.items():
            v = out.get(k, 0) + cy * c
            if v:
                out[k] = v
            else:
                out.pop(k, None)
        return out

    @staticmethod
    def f_scale(x, c):
        return {k: v * c for k, v in x.items()} if c else {}

    def form_of(self, b):
        return {b.id: 1}

    def fits(self, f):
        return nfields(f) <= TMAX and sum(abs(c) for c in f.values()) <= CSUM_MAX

    def normalise(self, f):
        """Keep forms encodable: materialise when they outgrow one LIN op."""
        if self.fits(f):
            return f
        return self.form_of(self.materialise(f))

    def materialise(self, f):
        """Form -> Base holding its canonical value."""
        if not f:
            return self.zero
        if len(f) == 1:
            (k, c), = f.items()
            if c == 1:
                return self.bases[k]
        if not self.fits(f):
            items = sorted(f.items())
            # (a) coefficients beyond the 6-bit field: c b = q (CMAX b) + rem b with CMAX b materialised once
            if not SPLIT_BIG and any(abs(c) > CMAX for _, c in items):
                g2 = {}
                for k, c in items:
                    if abs(c) > CMAX:
                        sg = 1 if c > 0 else -1
                        q, rem = divmod(abs(c), CMAX)
                        big = self.materialise({k: CMAX})
                        g2 = self.f_add(g2, {big.id: sg * q})
                        if rem:
                            g2 = self.f_add(g2, {k: sg * rem})
                    else:
                        g2 = self.f_add(g2, {k: c})
                return self.materialise(g2)
            # (b) too many terms / too large a coefficient sum: chunks in producer order, then the sum of the chunks
            chunks, cur, csum = [], {}, 0
            for k, c in items:
                while abs(c) > CSUM_MAX:  # (never in the pairing programs) peel full-budget pieces off a huge coefficient
                    if cur:
                        chunks.append(cur)
                    sg = 1 if c > 0 else -1
                    chunks.append({k: sg * CSUM_MAX})
                    cur, csum, c = {}, 0, c - sg * CSUM_MAX
                if nfields(cur) + nfields({k: c}) > TMAX or csum + abs(c) > CSUM_MAX:
                    chunks.append(cur)
                    cur, csum = {}, 0
                cur[k] = c
                csum += abs(c)
            chunks.append(cur)
            tot = {}
            for ch in chunks:
                tot = self.f_add(tot, self.form_of(self.materialise(ch)))
            return self.materialise(tot)
        key = tuple(sorted(f.items()))
        if key in self.lin_cache:
            return self.lin_cache[key]
        b = self._new("op")
        b.op = OP_LIN
        b.terms = key
        for k, _ in key:
            self.bases[k].users.append(b)
        self.lin_cache[key] = b
        return b

    def operand(self, f):
        """Form -> (s0, s1, neg): the value s0 + (-1)^neg s1 with both slots canonical, for a MUL operand."""
        if not f:
            return (self.zero, self.zero, 0)
        items = sorted(f.items())
        if len(items) == 1 and items[0][1] == 1:
            return (self.bases[items[0][0]], self.zero, 0)
        if len(items) == 2:
            (k0, c0), (k1, c1) = items
            if c0 == 1 and c1 in (1, -1):
                return (self.bases[k0], self.bases[k1], 1 if c1 < 0 else 0)
            if c1 == 1 and c0 == -1:
                return (self.bases[k1], self.bases[k0], 1)
        return (self.materialise(f), self.zero, 0)

    def mul(self, fa, fb):
        """Product of two forms -> form of one new base."""
        if not fa or not fb:
            return {}
        a, b = self.operand(fa), self.operand(fb)
        ka = (a[0].id, a[1].id, a[2])
        kb = (b[0].id, b[1].id, b[2])
        if kb < ka:
            a, b, ka, kb = b, a, kb, ka
        key = (ka, kb)
        if key in self.mul_cache:
            return self.form_of(self.mul_cache[key])
        n = self._new("op")
        n.op = OP_MUL
        n.src = (a[0], a[1], b[0], b[1])
        n.flags = a[2] | (b[2] << 1)
        for s in set(n.src):
            s.users.append(n)
        self.mul_cache[key] = n
        return self.form_of(n)

    def inv(self, f):
        a = self.materialise(f)
        n = self._new("op")
        n.op = OP_INV
        n.src = (a,)
        a.users.append(n)
        return self.form_of(n)


def lower(t, low, in_forms):
    """Walk the traced Fp2 DAG (vmgen.Tracer) in id order; val[node.id] = (form0, form1)."""
    A, S = low.f_add, low.f_scale
    N = low.normalise
    val = {}
    for n in t.nodes:
        if n.op is None:
            val[n.id] = in_forms[n.name]
            continue
        s = n.src
        v = lambda i: val[s[i].id] if s[i] is not None else ({}, {})
        op = n.op
        if op == g.LDC:
            c = g.CONST2[n.imm]
            r = (low.form_of(low.const(c[0])) if c[0] else {}, low.form_of(low.const(c[1])) if c[1] else {})
        elif op in (g.ADD, g.SUB, g.SUB2):
            a, b = v(0), v(2)
            sg = 1 if op == g.ADD else -1
            r = (A(a[0], b[0], sg), A(a[1], b[1], sg))
            if op == g.SUB2:
                c = v(3)
                r = (A(r[0], c[0], -1), A(r[1], c[1], -1))
        elif op == g.DBL:
            r = (S(v(0)[0], 2), S(v(0)[1], 2))
        elif op == g.TRIPLE:
            r = (S(v(0)[0], 3), S(v(0)[1], 3))
        elif op == g.NEG:
            r = (S(v(0)[0], -1), S(v(0)[1], -1))
        elif op == g.CONJ:
            r = (v(0)[0], S(v(0)[1], -1))
        elif op == g.MOV:
            r = v(0)
        elif op in (g.MULXI, g.ADDXI, g.SUBXI):
            b = v(0) if op == g.MULXI else v(2)
            if SHARE:  # both components of b enter both components of xi b
                b = tuple(f if (len(f) < SHARE_MIN or (len(f) <= 1 and all(c == 1 for c in f.values()))) else low.form_of(low.materialise(f)) for f in b)
            b = (N(b[0]), N(b[1]))
            x = (A(S(b[0], 9), b[1], -1), A(b[0], S(b[1], 9)))
            if op == g.MULXI:
                r = x
            else:
                a = v(0)
                sg = 1 if op == g.ADDXI else -1
                r = (A(a[0], x[0], sg), A(a[1], x[1], sg))
        elif op == g.HALF:
            h = low.form_of(low.const(INV2))
            r = (low.mul(v(0)[0], h), low.mul(v(0)[1], h))
        elif op in (g.MUL, g.SQR):
            a = v(0)
            if s[1] is not None:
                a = (A(a[0], v(1)[0]), A(a[1], v(1)[1]))
            if op == g.SQR:
                # complex squaring: (a0 + a1)(a0 - a1), 2 a0 a1
                a = (low.form_of(low.materialise(a[0])) if len(a[0]) != 1 or list(a[0].values())[0] != 1 else a[0],
                     low.form_of(low.materialise(a[1])) if len(a[1]) != 1 or list(a[1].values())[0] != 1 else a[1])
                r = (low.mul(A(a[0], a[1]), A(a[0], a[1], -1)), S(low.mul(a[0], a[1]), 2))
            else:
                b = v(2)
                if s[3] is not None:
                    b = (A(b[0], v(3)[0]), A(b[1], v(3)[1]))
                # Karatsuba over Fp: single-base operands so that the (a0 + a1) pre-addition rides in the MUL
                one = lambda f: f if (len(f) == 1 and list(f.values())[0] == 1) or not f else low.form_of(low.materialise(f))
                a, b = (one(a[0]), one(a[1])), (one(b[0]), one(b[1]))
                t0, t1 = low.mul(a[0], b[0]), low.mul(a[1], b[1])
                t2 = low.mul(A(a[0], a[1]), A(b[0], b[1]))
                r = (A(t0, t1, -1), A(A(t2, t0, -1), t1, -1))
                for k in (4, 5):  # fused subtrahends of the Karatsuba recombination
                    if s[k] is not None:
                        c = v(k)
                        r = (A(r[0], c[0], -1), A(r[1], c[1], -1))
        elif op == g.MULFP:
            k = v(2)[n.imm & 1]
            r = (low.mul(v(0)[0], k), low.mul(v(0)[1], k))
        elif op == g.MULCFP:
            k = low.form_of(low.const(g.CONSTFP[n.imm]))
            r = (low.mul(v(0)[0], k), low.mul(v(0)[1], k))
        elif op == g.MULC:
            c = g.CONST2[n.imm]
            a = v(0)
            one = lambda f: f if (len(f) == 1 and list(f.values())[0] == 1) or not f else low.form_of(low.materialise(f))
            a = (one(a[0]), one(a[1]))
            c0, c1, cs = low.form_of(low.const(c[0])), low.form_of(low.const(c[1])), low.form_of(low.const(c[0] + c[1]))
            t0, t1 = low.mul(a[0], c0), low.mul(a[1], c1)
            t2 = low.mul(A(a[0], a[1]), cs)
            r = (A(t0, t1, -1), A(A(t2, t0, -1), t1, -1))
        elif op == g.INV:
            a = v(0)
            nrm = A(low.mul(a[0], a[0]), low.mul(a[1], a[1]))
            ni = low.inv(nrm)
            r = (low.mul(a[0], ni), S(low.mul(a[1], ni), -1))
        else:
            raise ValueError(g.OPNAMES[op])
        r = (N(r[0]), N(r[1]))
        if SHARE and len(set(u.id for u in n.users)) > 1:
            # a value with several consumers is computed ONCE: inlining its form into each of them multiplies the work
            r = tuple(f if (len(f) < SHARE_MIN or (len(f) <= 1 and all(c == 1 for c in f.values()))) else low.form_of(low.materialise(f)) for f in r)
        val[n.id] = r
    return val


def inline_lins(low, outputs, tmax):
    """Collapse LIN -> LIN chains: a LIN term that is itself a LIN with no other consumer (or with at most two terms)
    is replaced by its own terms, as long as the merged op still fits one record.  Fewer dependent rounds; the
    absorbed node dies when nothing else reads it."""
    def users_of():
        cnt = {}
        for b in low.bases:
            if b.kind != "op":
                continue
            ds = b.src if b.op in (OP_MUL, OP_INV) else tuple(low.bases[k] for k, _ in b.terms)
            for d in set(x.id for x in ds):
                cnt[d] = cnt.get(d, 0) + 1
        for o_ in outputs:
            cnt[o_.id] = cnt.get(o_.id, 0) + 2  # outputs are never absorbed
        return cnt
    changed, total = True, 0
    while changed:
        changed = False
        cnt = users_of()
        for b in low.bases:
            if b.kind != "op" or b.op != OP_LIN:
                continue
            f = dict(b.terms)
            for k, c in list(f.items()):
                d = low.bases[k]
                if d.kind != "op" or d.op != OP_LIN or d is b:
                    continue
                if not (cnt.get(k, 0) == 1 or len(d.terms) <= DUP_MAX):
                    continue
                g2 = dict(f)
                del g2[k]
                g2 = low.f_add(g2, dict(d.terms), c)
                if nfields(g2) <= tmax and g2 and sum(abs(v) for v in g2.values()) <= CSUM_MAX:
                    f = g2
                    changed = True
                    total += 1
            b.terms = tuple(sorted(f.items()))
    return total


def schedule(low, outputs):
    """Critical-path list scheduling into class-uniform rounds of <= LANES ops."""
    live = set()
    stack = list(outputs)
    deps = lambda b: b.src if b.op in (OP_MUL, OP_INV) else tuple(low.bases[k] for k, _ in b.terms)
    while stack:
        b = stack.pop()
        if b.id in live:
            continue
        live.add(b.id)
        stack.extend(deps(b))
    ops = [b for b in low.bases if b.id in live and b.kind == "op"]
    users = {b.id: [] for b in low.bases}
    for b in ops:
        for d in set(deps(b)):
            users[d.id].append(b)
    # priority: longest weighted path to an output
    for b in reversed(ops):
        b.prio = COST[b.op] + max((u.prio for u in users[b.id]), default=0.0)
    indeg = {}
    ready = {OP_MUL: [], OP_LIN: [], OP_INV: []}
    for b in ops:
        d = sum(1 for x in set(deps(b)) if x.kind == "op")
        indeg[b.id] = d
        if d == 0:
            ready[b.op].append(b)
    rounds, done = [], 0
    while done < len(ops):
        best, key = None, None
        if POLICY == "lin_first" and ready[OP_LIN]:
            best = OP_LIN  # cheap rounds first: every LIN that can run makes more products ready for the next MUL round
        elif POLICY.startswith("fill") and ready[OP_LIN] and 0 < len(ready[OP_MUL]) < int(POLICY[4:]):
            best = OP_LIN  # an under-filled MUL round waits while cheap LIN rounds can still feed it
        else:
            for op, lst in ready.items():
                if not lst:
                    continue
                top = max(x.prio for x in lst)
                k = (top, min(len(lst), LANES))
                if key is None or k > key:
                    best, key = op, k
        lst = ready[best]
        lst.sort(key=lambda x: -x.prio)
        take, ready[best] = lst[:LANES], lst[LANES:]
        for b in take:
            b.round = len(rounds)
        rounds.append(take)
        done += len(take)
        for b in take:
            for u in users[b.id]:
                indeg[u.id] -= 1
                if indeg[u.id] == 0:
                    ready[u.op].append(u)
    return rounds, live


def allocate(low, rounds, live, outputs, pinned):
    """Round-granular linear scan.  pinned: bases with fixed slots (inputs, constants).  Outputs get their own slots."""
    deps = lambda b: b.src if b.op in (OP_MUL, OP_INV) else tuple(low.bases[k] for k, _ in b.terms)
    last = {}
    for r, ops in enumerate(rounds):
        for b in ops:
            for d in deps(b):
                last[d.id] = r
    for b in outputs:
        last[b.id] = len(rounds) + 1
    nxt = 0
    for b in pinned:
        b.slot = nxt
        nxt += 1
    free, expiring, peak = [], {}, nxt
    for r, ops in enumerate(rounds):
        for s in expiring.pop(r - 1, []):
            free.append(s)
        for b in ops:
            if free:
                b.slot = free.pop()
            else:
                b.slot = nxt
                nxt += 1
                peak = max(peak, nxt)
            expiring.setdefault(last.get(b.id, r), []).append(b.slot)
    return peak


def encode(low, rounds):
    """16 x u16 per op.  rec[0] = op | dst << 2 | (nterms or flags) << 12.
       MUL: rec[1..4] = s0, s1, s2, s3 ; flags bit0 = negate s1, bit1 = negate s3.
       LIN: rec[1 + i] = slot | (coef & 63) << 10.   INV: rec[1] = s0."""
    out = []
    for ops in rounds:
        for j in range(LANES):
            rec = [0] * 16
            if j < len(ops):
                b = ops[j]
                assert b.slot < 1024
                if b.op == OP_MUL:
                    rec[0] = OP_MUL | (b.slot << 2) | (b.flags << 12)
                    for i, s in enumerate(b.src):
                        rec[1 + i] = s.slot
                elif b.op == OP_INV:
                    rec[0] = OP_INV | (b.slot << 2)
                    rec[1] = b.src[0].slot
                else:
                    fields = []
                    for k, c in b.terms:  # a coefficient beyond the 6-bit field becomes several terms on the same slot
                        sg = 1 if c > 0 else -1
                        while abs(c) > CMAX:
                            fields.append((k, sg * CMAX))
                            c -= sg * CMAX
                        fields.append((k, c))
                    assert len(fields) <= 15
                    rec[0] = OP_LIN | (b.slot << 2) | (len(fields) << 12)
                    for i, (k, c) in enumerate(fields):
                        assert -32 <= c <= 31
                        rec[1 + i] = low.bases[k].slot | ((c & 63) << 10)
            out.extend(rec)
    return out


def evaluate(words, nslots, init):
    """Reference interpreter on Python ints (lock-step rounds: all lanes load before any lane stores)."""
    slots = [0] * nslots
    for s, v in init.items():
        slots[s] = v % P
    nrounds = len(words) // (16 * LANES)
    for r in range(nrounds):
        res = []
        for j in range(LANES):
            rec = words[(r * LANES + j) * 16:(r * LANES + j + 1) * 16]
            op, dst, x = rec[0] & 3, (rec[0] >> 2) & 1023, rec[0] >> 12
            if op == OP_NOP:
                continue
            if op == OP_MUL:
                a = slots[rec[1]] + (-slots[rec[2]] if x & 1 else slots[rec[2]])
                b = slots[rec[3]] + (-slots[rec[4]] if x & 2 else slots[rec[4]])
                v = a * b % P
            elif op == OP_INV:
                v = pow(slots[rec[1]], P - 2, P)
            else:
                v = 0
                for i in range(x):
                    c = rec[1 + i] >> 10
                    c = c - 64 if c >= 32 else c
                    v += c * slots[rec[1 + i] & 1023]
                v %= P
            res.append((dst, v))
        for d, v in res:
            slots[d] = v
    return slots


def build(name):
    """name: pair | miller | finalexp.  Returns (words, meta)."""
    saved = (g.park12, g.Tracer.park)
    g.park12 = lambda t, x: x            # everything lives in shared memory here
    g.Tracer.park = lambda self, a: a
    try:
        t = g.Tracer()
        low = Lowering()
        in_forms, pinned_in = {}, []
        if name == "miller2":  # product of two Miller loops with shared squarings (the BLS verification shape)
            Ps, Qs = [], []
            for j in range(2):
                names = ("P%d" % j, "Q%dx" % j, "Q%dy" % j)
                Pslot, Qx, Qy = (t.input(nm) for nm in names)
                for nm in names:
                    b0, b1 = low.input(nm + ".0"), low.input(nm + ".1")
                    pinned_in += [b0, b1]
                    in_forms[nm] = (low.form_of(b0), low.form_of(b1))
                Ps.append(Pslot); Qs.append((Qx, Qy))
            f = g.trace_miller_multi(t, Ps, Qs)
        elif name in ("pair", "miller"):
            Pslot, Qx, Qy = t.input("P"), t.input("Qx"), t.input("Qy")
            for nm in ("P", "Qx", "Qy"):
                b0, b1 = low.input(nm + ".0"), low.input(nm + ".1")
                pinned_in += [b0, b1]
                in_forms[nm] = (low.form_of(b0), low.form_of(b1))
            f = g.trace_miller(t, Pslot, (Qx, Qy))
            if name == "pair":
                f = g.trace_final_exp(t, f, park=False)
        else:
            gi = [t.input("f%d" % i) for i in range(6)]
            for i in range(6):
                b0, b1 = low.input("f%d.0" % i), low.input("f%d.1" % i)
                pinned_in += [b0, b1]
                in_forms["f%d" % i] = (low.form_of(b0), low.form_of(b1))
            f = g.trace_final_exp(t, ((gi[0], gi[1], gi[2]), (gi[3], gi[4], gi[5])), park=False)
        outs2 = g.flat12(f)
        val = lower(t, low, in_forms)
    finally:
        g.park12, g.Tracer.park = saved
    outputs = []
    for v in outs2:
        for comp in (0, 1):
            form = val[v.id][comp]
            b = low.materialise(form)
            if b.kind != "op" or b in outputs:  # outputs own distinct op slots: copy through a 1-term LIN
                c = low._new("op")
                c.op = OP_LIN
                c.terms = ((b.id, 1),)
                b.users.append(c)
                b = c
            outputs.append(b)
    tinline = int(TINLINE_ENV) if TINLINE_ENV is not None else int(os.environ.get("WVM_TINLINE_" + name.upper(), TINLINE_BY_PROGRAM[name]))
    if tinline:
        inline_lins(low, outputs, tinline)
    rounds, live = schedule(low, outputs)
    consts = [b for b in low.bases if b.kind == "const" and (b.id in live or b is low.zero)]
    pinned = [low.zero] + [b for b in consts if b is not low.zero] + pinned_in
    nslots = allocate(low, rounds, live, outputs, pinned)
    words = encode(low, rounds)
    nmul = sum(1 for r in rounds if r[0].op == OP_MUL)
    nlin = sum(1 for r in rounds if r[0].op == OP_LIN)
    meta = {"rounds": len(rounds), "mul_rounds": nmul, "lin_rounds": nlin, "inv_rounds": len(rounds) - nmul - nlin, "nslots": nslots,
            "mul_ops": sum(len(r) for r in rounds if r[0].op == OP_MUL), "lin_ops": sum(len(r) for r in rounds if r[0].op == OP_LIN),
            "lin_terms": sum(len(b.terms) for r in rounds for b in r if b.op == OP_LIN),
            "consts": [(b.slot, b.value_const) for b in pinned if b.kind == "const"],
            "in_slots": [b.slot for b in pinned_in], "out_slots": [b.slot for b in outputs]}
    return words, meta


def mont32(v):
    return ", ".join("0x%08xu" % (((v << 256) % P >> (32 * i)) & 0xFFFFFFFF) for i in range(8))


def emit(outdir):
    lines = ["// GENERATED by wvmgen.py -- do not edit.", "#pragma once", "namespace bn254 { namespace wvm {"]
    for name in ("miller", "miller2", "finalexp"):  # a pairing runs the Miller and final-exponentiation programs back to back (k_wvm.cu)
        words, meta = build(name)
        with open(os.path.join(outdir, "wvm_prog_%s.inc" % name), "w") as f:  # u32 words = two u16 fields, little-endian
            for i in range(0, len(words), 16):
                f.write(", ".join("0x%08xu" % (words[i + 2 * j] | (words[i + 2 * j + 1] << 16)) for j in range(8)) + ",\n")
        tag = name.upper()
        print(tag, {k: v for k, v in meta.items() if k not in ("consts",)})
        lines.append("// %s: %s" % (tag, {k: v for k, v in meta.items() if k not in ("consts", "in_slots", "out_slots")}))
        lines.append("static constexpr int %s_ROUNDS = %d;" % (tag, meta["rounds"]))
        lines.append("static constexpr int %s_NSLOTS = %d;" % (tag, meta["nslots"]))
        lines.append("static constexpr int %s_NCONST = %d;" % (tag, len(meta["consts"])))
        lines.append("BN_CONST uint16_t %s_CONST_SLOT[%d] = {%s};" % (tag, len(meta["consts"]), ", ".join(str(s) for s, _ in meta["consts"])))
        lines.append("BN_CONST Fp %s_CONST_VAL[%d] = {" % (tag, len(meta["consts"])))
        for _, v in meta["consts"]:
            lines.append("  {{%s}}," % mont32(v))
        lines.append("};")
        lines.append("BN_CONST uint16_t %s_IN[%d] = {%s};" % (tag, len(meta["in_slots"]), ", ".join(map(str, meta["in_slots"]))))
        lines.append("BN_CONST uint16_t %s_OUT[12] = {%s};" % (tag, ", ".join(map(str, meta["out_slots"]))))
    lines.append("} }  // namespace bn254::wvm")
    with open(os.path.join(outdir, "wvm_prog_meta.cuh"), "w") as f:
        f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    emit(HERE)
