"""Tower-VM program generator for the lane-group pairing kernels.

The CUDA interpreter (vm.cuh) executes straight-line programs of Fp2-level micro-ops over a per-pairing
file of 64-byte slots held in shared memory.  K lanes of a warp cooperate on one pairing: a program is a
sequence of ROUNDS, each round holds K independent micro-ops of the same opcode (one per lane).  This file

  1. traces the pairing formulas (same algebra as tower.cuh / pairing.cuh) into a DAG of micro-ops,
  2. list-schedules the DAG into K-wide rounds,
  3. allocates slots by liveness (values dead after their last consumer round are recycled),
  4. emits vm_prog_<name>_k<K>.inc (uint64 words) + vm_prog_meta.cuh,
  5. can EVALUATE a scheduled program on Python integers, which tests/ compare with the oracle.

Self-contained (no oracle import): the product never depends on oracle/.
Run from the repo root:  python gopairingbasedcryptography_b200/csrc/vmgen.py
"""
from __future__ import annotations

import os
import sys

X0 = 4965661367192848881
P = 36 * X0**4 + 36 * X0**3 + 24 * X0**2 + 6 * X0 + 1
R = 36 * X0**4 + 36 * X0**3 + 18 * X0**2 + 6 * X0 + 1

# ----------------------------------------------------------------------------------------- opcodes
(NOP, MUL, SQR, ADD, SUB, SUB2, DBL, NEG, CONJ, MULXI, HALF, MULFP, MULC, MULCFP, MOV, ADDXI, TRIPLE, INV, LDC,
 SUBXI) = range(20)
OPNAMES = ["NOP", "MUL", "SQR", "ADD", "SUB", "SUB2", "DBL", "NEG", "CONJ", "MULXI", "HALF", "MULFP", "MULC", "MULCFP",
           "MOV", "ADDXI", "TRIPLE", "INV", "LDC", "SUBXI"]
COST = {NOP: 0, MUL: 30, SQR: 20, ADD: 1, SUB: 1, SUB2: 1.5, DBL: 1, NEG: 1, CONJ: 1, MULXI: 2, HALF: 1, MULFP: 20,
        MULC: 30, MULCFP: 20, MOV: 0.5, ADDXI: 2.5, TRIPLE: 1.5, INV: 4000, LDC: 0.5, SUBXI: 2.5}
NONE = 0xFF


# ----------------------------------------------------------------------------------------- Fp2 on ints
def f2add(a, b):
    return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)


def f2sub(a, b):
    return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)


def f2mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def f2xi(a):
    return ((9 * a[0] - a[1]) % P, (a[0] + 9 * a[1]) % P)


def f2inv(a):
    n = pow(a[0] * a[0] + a[1] * a[1], -1, P) if (a[0] or a[1]) else 0
    return (a[0] * n % P, -a[1] * n % P)


def f2pow(a, e):
    out = (1, 0)
    while e:
        if e & 1:
            out = f2mul(out, a)
        a = f2mul(a, a)
        e >>= 1
    return out


XI = (9, 1)
INV2 = pow(2, -1, P)
GAMMA1 = [f2pow(XI, j * (P - 1) // 6) for j in range(6)]
GAMMA2 = [f2pow(XI, j * (P * P - 1) // 6)[0] for j in range(6)]
GAMMA3 = [f2pow(XI, j * (P**3 - 1) // 6) for j in range(6)]
TWIST_3B = tuple(3 * c % P for c in f2mul((3, 0), f2inv(XI)))
# Fp2 constant table (MULC / LDC imm) and Fp constant table (MULCFP imm)
CONST2 = [(0, 0), (1, 0), TWIST_3B] + GAMMA1[1:6] + GAMMA3[1:6]
C_ZERO, C_ONE, C_3B = 0, 1, 2
C_G1 = {j: 2 + j for j in range(1, 6)}  # GAMMA1[j] -> index
C_G3 = {j: 7 + j for j in range(1, 6)}
CONSTFP = GAMMA2[:]  # index j = GAMMA2[j]


def naf(k, width=2):
    out = []
    mod = 1 << width
    while k:
        if k & 1:
            d = k % mod
            if d >= mod // 2:
                d -= mod
            k -= d
        else:
            d = 0
        out.append(d)
        k >>= 1
    return out


# ----------------------------------------------------------------------------------------- tracing
class Node:
    __slots__ = ("id", "op", "src", "imm", "users", "round", "slot", "prio", "cold", "name")

    def __init__(self, id, op, src, imm):
        self.id, self.op, self.src, self.imm = id, op, src, imm
        self.users, self.round, self.slot, self.prio, self.cold, self.name = [], None, None, 0.0, False, None


class Tracer:
    """Builds the DAG.  Values are Node objects; inputs are nodes with op None (pre-loaded slots)."""

    def __init__(self):
        self.nodes = []
        self.cse = {}
        self.inputs = []
        self.outputs = []

    def _mk(self, op, src, imm=0):
        src = tuple(src) + (None,) * (6 - len(src))  # (a, a2, b, b2, c, e): MUL computes (a+a2)(b+b2) - c - e
        key = (op, tuple(s.id if s is not None else -1 for s in src), imm)
        if op not in (None,) and key in self.cse:
            return self.cse[key]
        n = Node(len(self.nodes), op, tuple(src), imm)
        self.nodes.append(n)
        for s in src:
            if s is not None:
                s.users.append(n)
        self.cse[key] = n
        return n

    def input(self, name, cold=False):
        n = Node(len(self.nodes), None, (), 0)
        n.name = name
        n.cold = cold
        self.nodes.append(n)
        self.inputs.append(n)
        return n

    def output(self, v, name):
        self.outputs.append((v, name))

    # micro-ops: src layout is always (a, a2, b, b2)
    def mul(self, a, b, a2=None, b2=None):
        if (b.id, -1 if b2 is None else b2.id) < (a.id, -1 if a2 is None else a2.id):
            a, a2, b, b2 = b, b2, a, a2
        return self._mk(MUL, (a, a2, b, b2))

    def sqr(self, a, a2=None):
        return self._mk(SQR, (a, a2, None, None))

    def add(self, a, b):
        if b.id < a.id:
            a, b = b, a
        return self._mk(ADD, (a, None, b, None))

    def sub(self, a, b):
        return self._mk(SUB, (a, None, b, None))

    def sub2(self, a, b, c):
        if c.id < b.id:
            b, c = c, b
        return self._mk(SUB2, (a, None, b, c))

    def dbl(self, a):
        return self._mk(DBL, (a, None, None, None))

    def triple(self, a):
        return self._mk(TRIPLE, (a, None, None, None))

    def neg(self, a):
        return self._mk(NEG, (a, None, None, None))

    def conj(self, a):
        return self._mk(CONJ, (a, None, None, None))

    def mulxi(self, a):
        return self._mk(MULXI, (a, None, None, None))

    def addxi(self, a, b):  # a + xi*b
        return self._mk(ADDXI, (a, None, b, None))

    def subxi(self, a, b):  # a - xi*b
        return self._mk(SUBXI, (a, None, b, None))

    def half(self, a):
        return self._mk(HALF, (a, None, None, None))

    def mulfp(self, a, s, comp):  # a * (component comp of slot s), an Fp scalar
        return self._mk(MULFP, (a, None, s, None), comp)

    def mulc(self, a, cidx):
        return self._mk(MULC, (a, None, None, None), cidx)

    def mulcfp(self, a, cidx):
        return self._mk(MULCFP, (a, None, None, None), cidx)

    def ldc(self, cidx):
        return self._mk(LDC, (None, None, None, None), cidx)

    def inv(self, a):
        return self._mk(INV, (a, None, None, None))

    def mov(self, a, cold=False):
        n = Node(len(self.nodes), MOV, (a, None, None, None, None, None), 0)  # never CSE'd: pins outputs / parks values
        n.cold = cold
        self.nodes.append(n)
        a.users.append(n)
        return n

    def park(self, a):
        """Copy into the COLD slot space (global memory): for values that stay idle for hundreds of rounds."""
        return self.mov(a, cold=True)


# ---- tower formulas on traced values (Fp6 = 3-tuple, Fp12 = (Fp6, Fp6)) ------------------------------
def fp6_add(t, x, y):
    return tuple(t.add(a, b) for a, b in zip(x, y))


def fp6_sub(t, x, y):
    return tuple(t.sub(a, b) for a, b in zip(x, y))


def fp6_neg(t, x):
    return tuple(t.neg(a) for a in x)


def fp6_mul(t, x, y):
    """Karatsuba: 6 products (pre-adds fused into the MUL operands), 6 recombination ops."""
    v0, v1, v2 = t.mul(x[0], y[0]), t.mul(x[1], y[1]), t.mul(x[2], y[2])
    t0 = t.mul(x[1], y[1], x[2], y[2])  # (x1+x2)(y1+y2)
    t1 = t.mul(x[0], y[0], x[1], y[1])
    t2 = t.mul(x[0], y[0], x[2], y[2])
    u0 = t.addxi(v0, t.sub2(t0, v1, v2))
    u1 = t.addxi(t.sub2(t1, v0, v1), v2)
    u2 = t.add(t.sub2(t2, v0, v2), v1)
    return (u0, u1, u2)


def fp6_mul_v(t, x):
    return (t.mulxi(x[2]), x[0], x[1])


def fp6_mul_fp2(t, x, k):
    return tuple(t.mul(a, k) for a in x)


def fp6_mul_01(t, x, c0, c1):
    a, b = t.mul(x[0], c0), t.mul(x[1], c1)
    r0 = t.addxi(a, t.sub(t.mul(x[1], c1, x[2], None), b))
    r2 = t.add(t.sub(t.mul(x[0], c0, x[2], None), a), b)
    r1 = t.sub2(t.mul(x[0], c0, x[1], c1), a, b)
    return (r0, r1, r2)


def fp6_inv(t, x):
    t0 = t.subxi(t.sqr(x[0]), t.mul(x[1], x[2]))
    t1 = t.sub(t.mulxi(t.sqr(x[2])), t.mul(x[0], x[1]))
    t2 = t.sub(t.sqr(x[1]), t.mul(x[0], x[2]))
    n = t.addxi(t.mul(x[0], t0), t.add(t.mul(x[2], t1), t.mul(x[1], t2)))
    ni = t.inv(n)
    return (t.mul(t0, ni), t.mul(t1, ni), t.mul(t2, ni))


def fp12_mul(t, x, y):
    a = fp6_mul(t, x[0], y[0])
    b = fp6_mul(t, x[1], y[1])
    s = fp6_add(t, x[0], x[1])
    u = fp6_add(t, y[0], y[1])
    c = fp6_mul(t, s, u)
    c1 = tuple(t.sub2(ci, ai, bi) for ci, ai, bi in zip(c, a, b))
    c0 = (t.addxi(a[0], b[2]), t.add(a[1], b[0]), t.add(a[2], b[1]))
    return (c0, c1)


def fp12_mul_conj(t, x, y):
    """x * conj(y) without materialising conj(y) = (y0, -y1): keeps the exponentiation table at one copy."""
    a = fp6_mul(t, x[0], y[0])
    b = fp6_mul(t, x[1], y[1])  # = -(x1 * conj(y)_1)
    s = fp6_add(t, x[0], x[1])
    u = fp6_sub(t, y[0], y[1])
    c = fp6_mul(t, s, u)
    c1 = tuple(t.add(t.sub(ci, ai), bi) for ci, ai, bi in zip(c, a, b))
    c0 = (t.subxi(a[0], b[2]), t.sub(a[1], b[0]), t.sub(a[2], b[1]))
    return (c0, c1)


def fp12_sqr(t, x):
    m = fp6_mul(t, x[0], x[1])
    s = fp6_add(t, x[0], x[1])
    u = (t.addxi(x[0][0], x[1][2]), t.add(x[0][1], x[1][0]), t.add(x[0][2], x[1][1]))  # c0 + v c1
    q = fp6_mul(t, s, u)
    # c0 = q - m - v m ; c1 = 2m
    vm = fp6_mul_v(t, m)
    c0 = tuple(t.sub2(qi, mi, vi) for qi, mi, vi in zip(q, m, vm))
    c1 = tuple(t.dbl(mi) for mi in m)
    return (c0, c1)


def fp12_conj(t, x):
    return (x[0], fp6_neg(t, x[1]))


def fp12_inv(t, x):
    n = fp6_sub(t, fp6_mul(t, x[0], x[0]), fp6_mul_v(t, fp6_mul(t, x[1], x[1])))
    ni = fp6_inv(t, n)
    return (fp6_mul(t, x[0], ni), fp6_neg(t, fp6_mul(t, x[1], ni)))


def w_basis(x):
    return [x[0][0], x[1][0], x[0][1], x[1][1], x[0][2], x[1][2]]


def from_w_basis(g):
    return ((g[0], g[2], g[4]), (g[1], g[3], g[5]))


def fp12_frob(t, x, k):
    g = w_basis(x)
    if k & 1:
        g = [t.conj(v) for v in g]
    out = [g[0]]
    for i in range(1, 6):
        if k == 2:
            out.append(t.mulcfp(g[i], i))
        else:
            out.append(t.mulc(g[i], (C_G1 if k == 1 else C_G3)[i]))
    return from_w_basis(out)


def fp4_sqr(t, a, b):
    a2, b2 = t.sqr(a), t.sqr(b)
    s = t.sqr(a, b)  # (a+b)^2
    return t.addxi(a2, b2), t.sub2(s, a2, b2)


def fp12_cyclo_sqr(t, x):
    g = w_basis(x)
    a0, a1 = fp4_sqr(t, g[0], g[3])
    b0, b1 = fp4_sqr(t, g[1], g[4])
    c0, c1 = fp4_sqr(t, g[2], g[5])
    c1 = t.mulxi(c1)

    def three_minus_two(a, gi):  # 3a - 2g
        return t.add(t.dbl(t.sub(a, gi)), a)

    def three_plus_two(a, gi):  # 3a + 2g
        return t.add(t.dbl(t.add(a, gi)), a)

    n = [None] * 6
    n[0] = three_minus_two(a0, g[0])
    n[3] = three_plus_two(a1, g[3])
    n[2] = three_minus_two(b0, g[2])
    n[5] = three_plus_two(b1, g[5])
    n[1] = three_plus_two(c1, g[1])
    n[4] = three_minus_two(c0, g[4])
    return from_w_basis(n)


def fp12_mul_034(t, z, l0, l1, l3):
    a = fp6_mul_fp2(t, z[0], l0)
    b = fp6_mul_01(t, z[1], l1, l3)
    s = fp6_add(t, z[0], z[1])
    c = fp6_mul_01(t, s, t.add(l0, l1), l3)
    c1 = tuple(t.sub2(ci, ai, bi) for ci, ai, bi in zip(c, a, b))
    c0 = (t.addxi(a[0], b[2]), t.add(a[1], b[0]), t.add(a[2], b[1]))
    return (c0, c1)


def fp12_cyclo_exp(t, x, digits, cold_table=True):
    """x^e, e as width-3 signed digits LSB first; x in the cyclotomic subgroup (inverse = conjugate)."""
    if cold_table and not all(v.cold for h in x for v in h):
        x = park12(t, x)
    x3 = fp12_mul(t, fp12_cyclo_sqr(t, x), x)
    if cold_table:
        x3 = park12(t, x3)
    tab = {1: x, 3: x3}
    acc = None
    for d in reversed(digits):
        if acc is not None:
            acc = fp12_cyclo_sqr(t, acc)
        if d:
            e = tab[abs(d)]
            if acc is None:
                acc = e if d > 0 else fp12_conj(t, e)
            else:
                acc = fp12_mul(t, acc, e) if d > 0 else fp12_mul_conj(t, acc, e)
    return acc


def g2_dbl_step(t, T):
    X, Y, Z = T
    A = t.half(t.mul(X, Y))
    B, C = t.sqr(Y), t.sqr(Z)
    E = t.mulc(C, C_3B)
    F = t.triple(E)
    G = t.half(t.add(B, F))
    H = t.sub2(t.sqr(Y, Z), B, C)
    J = t.sqr(X)
    X3 = t.mul(A, t.sub(B, F))
    Y3 = t.sub(t.sqr(G), t.triple(t.sqr(E)))
    Z3 = t.mul(B, H)
    return (X3, Y3, Z3), (t.neg(H), t.triple(J), t.sub(E, B))


def g2_add_step(t, T, Q, update=True):
    X, Y, Z = T
    O = t.sub(Y, t.mul(Q[1], Z))
    L = t.sub(X, t.mul(Q[0], Z))
    r2 = t.sub(t.mul(Q[0], O), t.mul(L, Q[1]))
    line = (L, t.neg(O), r2)
    if not update:
        return None, line
    C, D = t.sqr(O), t.sqr(L)
    E = t.mul(L, D)
    F = t.mul(Z, C)
    G = t.mul(X, D)
    H = t.sub(t.add(E, F), t.dbl(G))
    X3 = t.mul(L, H)
    Y3 = t.sub(t.mul(t.sub(G, H), O), t.mul(Y, E))
    Z3 = t.mul(E, Z)
    return (X3, Y3, Z3), line


def apply_line(t, f, Pslot, line):
    l0 = t.mulfp(line[0], Pslot, 1)  # r0 * yP
    l1 = t.mulfp(line[1], Pslot, 0)  # r1 * xP
    if f is None:  # f == 1: the product is the line itself
        zero = t.ldc(C_ZERO)
        return ((l0, zero, zero), (l1, line[2], zero))
    return fp12_mul_034(t, f, l0, l1, line[2])


def trace_miller(t, Pslot, Q, f=None):
    """Miller loop of one pair.  Pslot holds (xP, yP) as the two Fp halves of one slot; Q = (x, y)."""
    one = t.ldc(C_ONE)
    T = (Q[0], Q[1], one)
    negQ = (Q[0], t.neg(Q[1]))
    digits = naf(6 * X0 + 2)
    for i in range(len(digits) - 2, -1, -1):
        if f is not None:
            f = fp12_sqr(t, f)
        T, line = g2_dbl_step(t, T)
        f = apply_line(t, f, Pslot, line)
        if digits[i]:
            T, line = g2_add_step(t, T, Q if digits[i] > 0 else negQ)
            f = apply_line(t, f, Pslot, line)
    q1 = (t.mulc(t.conj(Q[0]), C_G1[2]), t.mulc(t.conj(Q[1]), C_G1[3]))
    q2 = (t.mulcfp(Q[0], 2), Q[1])
    T, line = g2_add_step(t, T, q1)
    f = apply_line(t, f, Pslot, line)
    _, line = g2_add_step(t, T, q2, update=False)
    f = apply_line(t, f, Pslot, line)
    return f


def trace_miller_multi(t, Pslots, Qs):
    """Miller product of several pairs with SHARED squarings of f (the value equals the product of the single-pair
    Miller values exactly: field arithmetic is exact).  The G2 steps of the pairs are independent of each other and of
    f^2, so a wide scheduler runs them side by side; only the sparse line multiplications chain through f."""
    one = t.ldc(C_ONE)
    Ts = [(Q[0], Q[1], one) for Q in Qs]
    negQs = [(Q[0], t.neg(Q[1])) for Q in Qs]
    digits = naf(6 * X0 + 2)
    f = None
    for i in range(len(digits) - 2, -1, -1):
        if f is not None:
            f = fp12_sqr(t, f)
        for j in range(len(Qs)):
            Ts[j], line = g2_dbl_step(t, Ts[j])
            f = apply_line(t, f, Pslots[j], line)
        if digits[i]:
            for j in range(len(Qs)):
                Ts[j], line = g2_add_step(t, Ts[j], Qs[j] if digits[i] > 0 else negQs[j])
                f = apply_line(t, f, Pslots[j], line)
    for j, Q in enumerate(Qs):
        q1 = (t.mulc(t.conj(Q[0]), C_G1[2]), t.mulc(t.conj(Q[1]), C_G1[3]))
        q2 = (t.mulcfp(Q[0], 2), Q[1])
        Ts[j], line = g2_add_step(t, Ts[j], q1)
        f = apply_line(t, f, Pslots[j], line)
        _, line = g2_add_step(t, Ts[j], q2, update=False)
        f = apply_line(t, f, Pslots[j], line)
    return f


def trace_final_exp(t, z, park=True):
    x0d = naf(X0, 3)
    f = fp12_mul(t, fp12_conj(t, z), fp12_inv(t, z))
    f = fp12_mul(t, fp12_frob(t, f, 2), f)
    expt = lambda v: fp12_cyclo_exp(t, v, x0d)
    fc = park12(t, f) if park else f  # f is needed again only at the very end
    f = fc
    t0 = fp12_cyclo_sqr(t, fp12_conj(t, expt(f)))
    t1 = fp12_mul(t, t0, fp12_cyclo_sqr(t, t0))
    t0 = park12(t, t0) if park else t0
    t2 = fp12_conj(t, expt(t1))
    t1 = fp12_mul(t, t2, fp12_conj(t, t1))
    t3 = fp12_cyclo_sqr(t, t2)
    if park:
        t1, t2 = park12(t, t1), park12(t, t2)
    t4 = fp12_mul(t, t1, expt(t3))
    f = fc
    t3 = fp12_mul(t, t0, t4)
    t0 = fp12_mul(t, f, fp12_mul(t, t2, t4))
    t0 = fp12_mul(t, fp12_frob(t, t3, 1), t0)
    t0 = fp12_mul(t, fp12_frob(t, t4, 2), t0)
    t2 = fp12_frob(t, fp12_mul(t, fp12_conj(t, f), t3), 3)
    return fp12_mul(t, t2, t0)


def park12(t, x):
    return tuple(tuple(t.park(v) for v in h) for h in x)


def flat12(x):
    return [x[0][0], x[0][1], x[0][2], x[1][0], x[1][1], x[1][2]]  # gnark memory order


# ----------------------------------------------------------------------------------------- scheduling
FUSE_MUL_SUB = os.environ.get("VM_FUSE", "1") == "1"


def fuse_mul_sub(t):
    """Peephole on the DAG: SUB2(m, x, y) / SUB(m, x) where m is a MUL used nowhere else becomes ONE MUL with the
    subtraction(s) folded in (the Karatsuba cross terms (a1+a2)(b1+b2) - v1 - v2)."""
    outs = set(v.id for v, _ in t.outputs)
    fused = 0
    for n in t.nodes:
        if n.op not in (SUB, SUB2):
            continue
        m = n.src[0]
        if m is None or m.op != MUL or m.id in outs or len(m.users) != 1 or m.src[4] is not None:
            continue
        c = n.src[2]
        e = n.src[3] if n.op == SUB2 else None
        if c is m or e is m:
            continue
        # rewrite n in place as the fused MUL; m dies (no users)
        for s_ in n.src:
            if s_ is not None and n in s_.users:
                s_.users.remove(n)
        n.op = MUL
        n.src = (m.src[0], m.src[1], m.src[2], m.src[3], c, e)
        for s_ in n.src:
            if s_ is not None:
                s_.users.append(n)
        m.users = []
        fused += 1
    return fused


def schedule(t, K, window=400):
    """List-schedule into rounds of <= K ops of one opcode.  Returns list of rounds (lists of nodes).
    Ops are taken in trace order (the natural depth-first order of the formulas, which keeps few values
    live) but any ready op within `window` ids of the oldest unscheduled op may fill a round."""
    nodes = [n for n in t.nodes if n.op is not None]
    # drop dead code (not reachable from outputs)
    live = set()
    stack = [v for v, _ in t.outputs]
    while stack:
        n = stack.pop()
        if n.id in live:
            continue
        live.add(n.id)
        for s in n.src:
            if s is not None:
                stack.append(s)
    nodes = [n for n in nodes if n.id in live]
    for n in t.nodes:
        n.users = [u for u in n.users if u.id in live]
    # priority = trace order (older first)
    for n in nodes:
        n.prio = -float(n.id)
    order = [n.id for n in nodes]
    scheduled = set()
    ptr = 0
    indeg = {}
    ready = {}
    for n in nodes:
        d = sum(1 for s in set(x for x in n.src if x is not None) if s.op is not None)
        indeg[n.id] = d
        if d == 0:
            ready.setdefault(n.op, []).append(n)
    rounds = []
    done = 0
    total = len(nodes)
    while done < total:
        while order[ptr] in scheduled:
            ptr += 1
        limit = order[ptr] + window
        # choose the opcode class: prefer full rounds, then the oldest op
        best_op, best_key = None, None
        for op, lst in ready.items():
            elig = [x for x in lst if x.id <= limit]
            if not elig:
                continue
            fill = min(len(elig), K) / K
            top = max(x.prio for x in elig)
            key = (fill >= 1.0, top, fill)
            if best_key is None or key > best_key:
                best_op, best_key = op, key
        lst = ready[best_op]
        lst.sort(key=lambda x: -x.prio)
        take = [x for x in lst if x.id <= limit][:K]
        tk = set(x.id for x in take)
        ready[best_op] = [x for x in lst if x.id not in tk]
        if MIX_ALU and len(take) < K and best_op in ALU_SET:
            # top up an under-filled add-type round with other eligible add-type ops (cheap divergence)
            extra = sorted((x for op2, l2 in ready.items() if op2 in ALU_SET and op2 != best_op for x in l2 if x.id <= limit),
                           key=lambda x: -x.prio)[: K - len(take)]
            for x in extra:
                ready[x.op].remove(x)
                take.append(x)
                tk.add(x.id)
        scheduled.update(tk)
        r = len(rounds)
        for n in take:
            n.round = r
        rounds.append(take)
        done += len(take)
        for n in take:
            for u in n.users:
                pass
        newly = []
        for n in take:
            for u in set(n.users):
                indeg[u.id] -= 1
                if indeg[u.id] == 0:
                    newly.append(u)
        for u in newly:
            ready.setdefault(u.op, []).append(u)
    return rounds


COLD_BASE = 160  # slot ids >= COLD_BASE live in global memory (parked values)


ALU_SET = {ADD, SUB, SUB2, DBL, NEG, CONJ, MULXI, HALF, MOV, ADDXI, TRIPLE, LDC, SUBXI}
MIX_ALU = os.environ.get("VM_MIX", "0") == "1"  # measured slower: divergent add-type rounds cost more than the fill they gain


def auto_cold(t, rounds, min_lifetime):
    """Values that stay live for more than `min_lifetime` rounds (exponentiation tables and their CSE'd
    Karatsuba sums, parked partial results) are produced straight into the cold space."""
    defr, last = {}, {}
    for r, ops in enumerate(rounds):
        for n in ops:
            defr[n.id] = r
            for s in n.src:
                if s is not None:
                    last[s.id] = r
    outs = set(v.id for v, _ in t.outputs)
    count = 0
    for ops in rounds:
        for n in ops:
            if not n.cold and n.id not in outs and last.get(n.id, defr[n.id]) - defr[n.id] > min_lifetime:
                n.cold = True
                count += 1
    return count


def allocate(t, rounds, nslots_max=COLD_BASE):
    """Linear-scan slot allocation at round granularity.  Inputs keep fixed slots (pre-loaded); outputs are
    MOVed into fixed slots at the end by the tracer (pinned)."""
    last_use = {}
    for r, ops in enumerate(rounds):
        for n in ops:
            for s in n.src:
                if s is not None:
                    last_use[s.id] = r
    for v, _ in t.outputs:
        last_use[v.id] = len(rounds) + 1
    free = []
    cold_free, cold_next = [], COLD_BASE
    next_slot = 0
    for n in t.inputs:
        if n.cold:
            n.slot = cold_next
            cold_next += 1
        else:
            n.slot = next_slot
            next_slot += 1
    expiring = {}
    for n in t.inputs:
        expiring.setdefault(last_use.get(n.id, -1), []).append(n.slot)
    peak = next_slot
    for r, ops in enumerate(rounds):
        # slots whose last use was in an EARLIER round are free now (no same-round recycling)
        for s in expiring.pop(r - 1, []):
            (cold_free if s >= COLD_BASE else free).append(s)
        for n in ops:
            if n.cold:
                if cold_free:
                    n.slot = cold_free.pop()
                else:
                    n.slot = cold_next
                    cold_next += 1
                    if cold_next > 254:
                        raise RuntimeError("cold slot space exhausted")
            elif free:
                n.slot = free.pop()
            else:
                n.slot = next_slot
                next_slot += 1
                peak = max(peak, next_slot)
            lu = last_use.get(n.id, r)
            expiring.setdefault(lu, []).append(n.slot)
    if peak > nslots_max:
        raise RuntimeError("slot file too large: %d" % peak)
    return peak, cold_next - COLD_BASE


def encode(rounds, K):
    words = []
    for ops in rounds:
        for j in range(K):
            if j < len(ops):
                n = ops[j]
                f = [s.slot if s is not None else NONE for s in n.src]
                if n.op == MUL:  # bits 48-55 / 56-63: slots subtracted from the product (NONE = absent)
                    w = n.op | (n.slot << 8) | (f[0] << 16) | (f[1] << 24) | (f[2] << 32) | (f[3] << 40) | (f[4] << 48) | (f[5] << 56)
                else:
                    w = n.op | (n.slot << 8) | (f[0] << 16) | (f[1] << 24) | (f[2] << 32) | (f[3] << 40) | (n.imm << 48)
            else:
                w = NOP
            words.append(w)
    return words


def evaluate(words, K, nslots, slots_init):
    """Reference interpreter on Python ints: slots_init = {slot: (a0, a1)}.  Returns the slot file."""
    slots = [(0, 0)] * nslots
    for s, v in slots_init.items():
        slots[s] = v
    for r in range(len(words) // K):
        results = []
        for j in range(K):
            w = words[r * K + j]
            op, d = w & 0xFF, (w >> 8) & 0xFF
            a, a2, b, b2, imm = (w >> 16) & 0xFF, (w >> 24) & 0xFF, (w >> 32) & 0xFF, (w >> 40) & 0xFF, (w >> 48) & 0xFF
            if op == NOP:
                continue
            A = slots[a] if a != NONE else None
            if a2 != NONE:
                A = f2add(A, slots[a2])
            B = slots[b] if b != NONE else None
            if op in (MUL,) and b2 != NONE:
                B = f2add(B, slots[b2])
            if op == MUL:
                v = f2mul(A, B)
                c, e = (w >> 48) & 0xFF, (w >> 56) & 0xFF
                if c != NONE:
                    v = f2sub(v, slots[c])
                if e != NONE:
                    v = f2sub(v, slots[e])
            elif op == SQR:
                v = f2mul(A, A)
            elif op == ADD:
                v = f2add(A, B)
            elif op == SUB:
                v = f2sub(A, B)
            elif op == SUB2:
                v = f2sub(f2sub(A, B), slots[b2])
            elif op == DBL:
                v = f2add(A, A)
            elif op == TRIPLE:
                v = f2add(f2add(A, A), A)
            elif op == NEG:
                v = f2sub((0, 0), A)
            elif op == CONJ:
                v = (A[0], -A[1] % P)
            elif op == MULXI:
                v = f2xi(A)
            elif op == ADDXI:
                v = f2add(A, f2xi(B))
            elif op == SUBXI:
                v = f2sub(A, f2xi(B))
            elif op == HALF:
                v = (A[0] * INV2 % P, A[1] * INV2 % P)
            elif op == MULFP:
                k = B[imm & 1]
                v = (A[0] * k % P, A[1] * k % P)
            elif op == MULC:
                v = f2mul(A, CONST2[imm])
            elif op == MULCFP:
                v = (A[0] * CONSTFP[imm] % P, A[1] * CONSTFP[imm] % P)
            elif op == MOV:
                v = A
            elif op == LDC:
                v = CONST2[imm]
            elif op == INV:
                v = f2inv(A)
            else:
                raise ValueError(op)
            results.append((d, v))
        for d, v in results:  # all lanes load before any lane stores
            slots[d] = v
    return slots


# ----------------------------------------------------------------------------------------- programs
def build_pair_program(K, with_miller=True, with_final_exp=True, window=400, park=True, cold_lifetime=0):
    """Inputs: slot 0 = (xP, yP), slot 1 = Q.x, slot 2 = Q.y (pair) or slots 0..5 = f (final exp only).
    Outputs: 6 slots holding the result in gnark memory order."""
    t = Tracer()
    if with_miller:
        Pslot, Qx, Qy = t.input("P", True), t.input("Qx", True), t.input("Qy", True)
        f = trace_miller(t, Pslot, (Qx, Qy))
    else:
        g = [t.input("f%d" % i) for i in range(6)]
        f = ((g[0], g[1], g[2]), (g[3], g[4], g[5]))
    if with_final_exp:
        f = trace_final_exp(t, f, park)
    outs = [t.mov(v) for v in flat12(f)]
    for i, v in enumerate(outs):
        t.output(v, "out%d" % i)
    if FUSE_MUL_SUB:
        fuse_mul_sub(t)
    rounds = schedule(t, K, window)
    if cold_lifetime:
        auto_cold(t, rounds, cold_lifetime)
    nslots, ncold = allocate(t, rounds)
    words = encode(rounds, K)
    meta = {"K": K, "rounds": len(rounds), "nslots": nslots, "ncold": ncold, "window": window, "in_slots": [n.slot for n in t.inputs],
            "out_slots": [v.slot for v, _ in t.outputs],
            "ops": sum(len(r) for r in rounds), "mul_rounds": sum(1 for r in rounds if r[0].op in (MUL, SQR, MULFP, MULC, MULCFP)),
            "fill": sum(len(r) for r in rounds) / (K * len(rounds))}
    hist = {}
    for r in rounds:
        for n in r:
            hist[OPNAMES[n.op]] = hist.get(OPNAMES[n.op], 0) + 1
    meta["hist"] = hist
    return words, meta


WINDOW = {1: 0, 2: 20, 3: 30, 4: 40, 6: 60}
COLD_LIFETIME = int(os.environ.get("VM_COLD_LIFETIME", "60"))  # rounds; longer-lived values go to the cold (global, L2-resident) slot space


def mont32(v):
    return ", ".join("0x%08xu" % (((v << 256) % P >> (32 * i)) & 0xFFFFFFFF) for i in range(8))


def emit(outdir, Ks=(3,)):
    lines = ["// GENERATED by vmgen.py -- do not edit.", "#pragma once", "namespace bn254 { namespace vm {"]
    lines.append("static constexpr int N_CONST2 = %d;" % len(CONST2))
    lines.append("BN_CONST Fp2 VM_CONST2[%d] = {" % len(CONST2))
    for c in CONST2:
        lines.append("  {{{%s}}, {{%s}}}," % (mont32(c[0]), mont32(c[1])))
    lines.append("};")
    lines.append("BN_CONST Fp VM_CONSTFP[%d] = {" % len(CONSTFP))
    for c in CONSTFP:
        lines.append("  {{%s}}," % mont32(c))
    lines.append("};")
    for name, kw in (("pair", {}), ("miller", {"with_final_exp": False}), ("finalexp", {"with_miller": False})):
        for K in Ks:
            words, meta = build_pair_program(K, window=WINDOW.get(K, 30), cold_lifetime=COLD_LIFETIME, **kw)
            inc = "vm_prog_%s_k%d.inc" % (name, K)
            with open(os.path.join(outdir, inc), "w") as f:
                for i in range(0, len(words), 4):
                    f.write(", ".join("0x%016xull" % w for w in words[i:i + 4]) + ",\n")
            tag = "%s_K%d" % (name.upper(), K)
            lines.append("// %s: %s" % (tag, {k: v for k, v in meta.items() if k != "hist"}))
            lines.append("//   op histogram: %s" % meta["hist"])
            lines.append("static constexpr int %s_ROUNDS = %d;" % (tag, meta["rounds"]))
            lines.append("static constexpr int %s_NSLOTS = %d;" % (tag, meta["nslots"]))
            lines.append("static constexpr int %s_NCOLD = %d;" % (tag, meta["ncold"]))
            lines.append("BN_CONST int %s_IN[%d] = {%s};" % (tag, len(meta["in_slots"]), ", ".join(map(str, meta["in_slots"]))))
            lines.append("BN_CONST int %s_OUT[6] = {%s};" % (tag, ", ".join(map(str, meta["out_slots"]))))
            print(tag, {k: v for k, v in meta.items() if k != "hist"})
    lines.append("} }  // namespace bn254::vm")
    with open(os.path.join(outdir, "vm_prog_meta.cuh"), "w") as f:
        f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    Ks = tuple(int(x) for x in sys.argv[1:]) or (3,)  # other lane widths (1, 2, 4, 6) regenerate on demand
    emit(here, Ks)
