"""Definitional big-integer oracle for the BN254 hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this package.  The product (the CUDA engine
behind ``include/bn254_b200.h``) never calls into it.

PARITY UNPINNED AGAINST GNARK, with one external anchor: the arithmetic the reference
executes lives in the third-party Go module ``github.com/consensys/gnark-crypto v0.19.0``
(/root/reference/go.mod:5), which is not on disk, and no Go toolchain exists in this
image.  The reference's own tests hold no golden vectors for Pair / ScalarMultiplication /
GT.Exp.  What IS pinned from outside the repository: G1 Add and
ScalarMultiplication by the EIP-196 precompile vectors (exact outputs), and the EIP-197
pairing-check known-answer vector 'jeff1' of go-ethereum's bn256Pairing test set
(tests/golden/eip197_pairing_check.json, tests/test_external_kat.py) holds -- it fixes
the curve, the twist, the G2 generator, the group laws and the Miller loop + final
exponentiation as a non-degenerate bilinear map.  A check against 1 cannot see the cofactor
of the final exponent, so the exact GT bytes (gnark's cofactor s below) remain unpinned.
This file restates the *published* definitions:

* optimal-ate Miller function  f_{6x+2,Q}(P) * l_{[6x+2]Q,pi(Q)}(P) * l_{[6x+2]Q+pi(Q),-pi^2(Q)}(P)
  evaluated with textbook affine chord/tangent lines (no projective formulas, no
  addition chains), and
* gnark's final exponent  d' = s*(p^12-1)/r  with  s = 2*x0*(6*x0^2+3*x0+1)
  (the "hard part up to permutation" of Fuentes-Castaneda et al. / Duquesne-Ghammam
  that gnark's ``FinalExponentiation`` documents), evaluated as a literal ``pow``.

Call sites this mirrors (reference file:line):
  bn254.Pair            access/tree/access_tree_node.go:106,110 ; cpabe/bsw07/bsw07_cpabe.go:184
  bn254.PairingCheck    signature/bls01_signature/bls_signature.go:81-84
  ScalarMultiplication  signature/bls01_signature/bls_signature.go:45,63
  GT.Exp                access/tree/access_tree_node.go:156 ; ibe/waters05_ibe/waters05_ibe.go:219
  GT.Mul/Div            access/tree/access_tree_node.go:114,157 ; cpabe/bsw07/bsw07_cpabe.go:189-190

Representation: Fp ints; Fp2 = (a0,a1) = a0+a1*u, u^2=-1; Fp6 = (b0,b1,b2) over v,
v^3 = xi = 9+u; Fp12 = (c0,c1) over w, w^2 = v.  This is gnark's E2/E6/E12 order.
Points are affine tuples (x, y) or None for infinity.
"""
from __future__ import annotations

P = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
X0 = 4965661367192848881
ATE_LOOP = 6 * X0 + 2
assert P == 36 * X0**4 + 36 * X0**3 + 24 * X0**2 + 6 * X0 + 1
assert R == 36 * X0**4 + 36 * X0**3 + 18 * X0**2 + 6 * X0 + 1
# gnark's final exponent: (p^12-1)/r times the cofactor s of its hard part
FE_COFACTOR = 2 * X0 * (6 * X0 * X0 + 3 * X0 + 1)
FE_EXPONENT = FE_COFACTOR * ((P**12 - 1) // R)
LAMBDA_GLV = 36 * X0**3 + 18 * X0**2 + 6 * X0 + 1  # lambda^2+lambda+1 = 0 mod r
MONT_R = 1 << 256

# ----------------------------------------------------------------------------- Fp2
FP2_ZERO = (0, 0)
FP2_ONE = (1, 0)
XI = (9, 1)


def fp2_add(a, b):
    return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)


def fp2_sub(a, b):
    return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)


def fp2_neg(a):
    return (-a[0] % P, -a[1] % P)


def fp2_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def fp2_sqr(a):
    return fp2_mul(a, a)


def fp2_scale(a, k):
    return (a[0] * k % P, a[1] * k % P)


def fp2_conj(a):
    return (a[0], -a[1] % P)


def fp2_inv(a):
    n = pow(a[0] * a[0] + a[1] * a[1], -1, P)
    return (a[0] * n % P, -a[1] * n % P)


def fp2_mul_xi(a):
    return ((9 * a[0] - a[1]) % P, (a[0] + 9 * a[1]) % P)


def fp2_pow(a, e):
    out = FP2_ONE
    while e:
        if e & 1:
            out = fp2_mul(out, a)
        a = fp2_sqr(a)
        e >>= 1
    return out


# ----------------------------------------------------------------------------- Fp6
FP6_ZERO = (FP2_ZERO, FP2_ZERO, FP2_ZERO)
FP6_ONE = (FP2_ONE, FP2_ZERO, FP2_ZERO)


def fp6_add(a, b):
    return tuple(fp2_add(x, y) for x, y in zip(a, b))


def fp6_sub(a, b):
    return tuple(fp2_sub(x, y) for x, y in zip(a, b))


def fp6_neg(a):
    return tuple(fp2_neg(x) for x in a)


def fp6_mul(a, b):
    a0, a1, a2 = a
    b0, b1, b2 = b
    r0 = fp2_add(fp2_mul(a0, b0), fp2_mul_xi(fp2_add(fp2_mul(a1, b2), fp2_mul(a2, b1))))
    r1 = fp2_add(fp2_add(fp2_mul(a0, b1), fp2_mul(a1, b0)), fp2_mul_xi(fp2_mul(a2, b2)))
    r2 = fp2_add(fp2_add(fp2_mul(a0, b2), fp2_mul(a1, b1)), fp2_mul(a2, b0))
    return (r0, r1, r2)


def fp6_mul_v(a):
    return (fp2_mul_xi(a[2]), a[0], a[1])


def fp6_inv(a):
    a0, a1, a2 = a
    t0 = fp2_sub(fp2_sqr(a0), fp2_mul_xi(fp2_mul(a1, a2)))
    t1 = fp2_sub(fp2_mul_xi(fp2_sqr(a2)), fp2_mul(a0, a1))
    t2 = fp2_sub(fp2_sqr(a1), fp2_mul(a0, a2))
    n = fp2_add(fp2_mul(a0, t0), fp2_mul_xi(fp2_add(fp2_mul(a2, t1), fp2_mul(a1, t2))))
    ni = fp2_inv(n)
    return (fp2_mul(t0, ni), fp2_mul(t1, ni), fp2_mul(t2, ni))


# ----------------------------------------------------------------------------- Fp12
FP12_ZERO = (FP6_ZERO, FP6_ZERO)
FP12_ONE = (FP6_ONE, FP6_ZERO)


def fp12_mul(a, b):
    a0, a1 = a
    b0, b1 = b
    return (
        fp6_add(fp6_mul(a0, b0), fp6_mul_v(fp6_mul(a1, b1))),
        fp6_add(fp6_mul(a0, b1), fp6_mul(a1, b0)),
    )


def fp12_sqr(a):
    return fp12_mul(a, a)


def fp12_conj(a):
    return (a[0], fp6_neg(a[1]))


def fp12_inv(a):
    a0, a1 = a
    n = fp6_sub(fp6_mul(a0, a0), fp6_mul_v(fp6_mul(a1, a1)))
    ni = fp6_inv(n)
    return (fp6_mul(a0, ni), fp6_neg(fp6_mul(a1, ni)))


def fp12_pow(a, e):
    """Left-to-right binary exponentiation; e >= 0."""
    if e == 0:
        return FP12_ONE
    out = a
    for bit in bin(e)[3:]:
        out = fp12_sqr(out)
        if bit == "1":
            out = fp12_mul(out, a)
    return out


def fp12_is_zero(a):
    return a == FP12_ZERO


def fp12_coeffs(a):
    """The 12 Fp coefficients in gnark memory order C0.B0.A0, C0.B0.A1, C0.B1.A0, ..."""
    return [a[i][j][k] for i in range(2) for j in range(3) for k in range(2)]


def fp12_from_coeffs(c):
    c = [x % P for x in c]
    return tuple(tuple((c[i * 6 + j * 2], c[i * 6 + j * 2 + 1]) for j in range(3)) for i in range(2))


def fp12_from_w_basis(g):
    """Build from sum g[i]*w^i, g[i] in Fp2 (w^2 = v)."""
    return ((g[0], g[2], g[4]), (g[1], g[3], g[5]))


# Frobenius constants gamma_{k,j} = xi^(j*(p^k-1)/6)
GAMMA1 = [fp2_pow(XI, j * (P - 1) // 6) for j in range(6)]
GAMMA2 = [fp2_pow(XI, j * (P * P - 1) // 6) for j in range(6)]
GAMMA3 = [fp2_pow(XI, j * (P**3 - 1) // 6) for j in range(6)]
assert all(g[1] == 0 for g in GAMMA2)


def fp12_frobenius(a, k=1):
    """a^(p^k) for k in 1,2,3 using the w-basis: g_i -> conj^k(g_i) * gamma_{k,i}."""
    gam = {1: GAMMA1, 2: GAMMA2, 3: GAMMA3}[k]
    g = [a[0][0], a[1][0], a[0][1], a[1][1], a[0][2], a[1][2]]
    out = []
    for i, gi in enumerate(g):
        if k & 1:
            gi = fp2_conj(gi)
        out.append(fp2_mul(gi, gam[i]))
    return fp12_from_w_basis(out)


# ----------------------------------------------------------------------------- curves
B1 = 3
B2 = fp2_mul((3, 0), fp2_inv(XI))  # twist coefficient 3/(9+u)
G1_GEN = (1, 2)
G2_GEN = (
    (
        10857046999023057135944570762232829481370756359578518086990519993285655852781,
        11559732032986387107991004021392285783925812861821192530917403151452391805634,
    ),
    (
        8495653923123431417604973247489272438418190587263600148770280649306958101930,
        4082367875863433681332203403145435568316851327593401208105741076214120093531,
    ),
)


def g1_on_curve(pt):
    if pt is None:
        return True
    x, y = pt
    return (y * y - x * x * x - B1) % P == 0


def g2_on_curve(pt):
    if pt is None:
        return True
    x, y = pt
    return fp2_sub(fp2_sqr(y), fp2_add(fp2_mul(fp2_sqr(x), x), B2)) == FP2_ZERO


def g1_neg(pt):
    return None if pt is None else (pt[0], -pt[1] % P)


def g1_add(a, b):
    """Affine group law with gnark's G1Affine.Add semantics (infinity, doubling, P + -P)."""
    if a is None:
        return b
    if b is None:
        return a
    if a[0] == b[0]:
        if (a[1] + b[1]) % P == 0:
            return None
        lam = 3 * a[0] * a[0] * pow(2 * a[1], -1, P) % P
    else:
        lam = (b[1] - a[1]) * pow(b[0] - a[0], -1, P) % P
    x3 = (lam * lam - a[0] - b[0]) % P
    return (x3, (lam * (a[0] - x3) - a[1]) % P)


def g1_mul(pt, k):
    """[k]pt for any integer k (negative k negates, as gnark's big.Int path does)."""
    if k < 0:
        return g1_mul(g1_neg(pt), -k)
    k %= R
    acc = None
    while k:
        if k & 1:
            acc = g1_add(acc, pt)
        pt = g1_add(pt, pt)
        k >>= 1
    return acc


def g2_neg(pt):
    return None if pt is None else (pt[0], fp2_neg(pt[1]))


def g2_add(a, b):
    if a is None:
        return b
    if b is None:
        return a
    if a[0] == b[0]:
        if fp2_add(a[1], b[1]) == FP2_ZERO:
            return None
        lam = fp2_mul(fp2_scale(fp2_sqr(a[0]), 3), fp2_inv(fp2_scale(a[1], 2)))
    else:
        lam = fp2_mul(fp2_sub(b[1], a[1]), fp2_inv(fp2_sub(b[0], a[0])))
    x3 = fp2_sub(fp2_sub(fp2_sqr(lam), a[0]), b[0])
    return (x3, fp2_sub(fp2_mul(lam, fp2_sub(a[0], x3)), a[1]))


def g2_mul(pt, k):
    if k < 0:
        return g2_mul(g2_neg(pt), -k)
    k %= R
    acc = None
    while k:
        if k & 1:
            acc = g2_add(acc, pt)
        pt = g2_add(pt, pt)
        k >>= 1
    return acc


def g2_frobenius(pt, k=1):
    """Untwist-Frobenius-twist endomorphism pi^k on E'(Fp2), k in {1,2}."""
    if pt is None:
        return None
    x, y = pt
    if k == 1:
        return (fp2_mul(fp2_conj(x), GAMMA1[2]), fp2_mul(fp2_conj(y), GAMMA1[3]))
    return (fp2_mul(x, GAMMA2[2]), fp2_mul(y, GAMMA2[3]))


# ----------------------------------------------------------------------------- pairing
def _line(T, S, Pt):
    """Chord (T != +-S) or tangent (T == S) through twist points, evaluated at the G1
    point Pt through the untwist (x',y') -> (x' w^2, y' w^3):
        l = yP - lam*xP*w + (lam*xT - yT)*w^3        
    Returns (line as Fp12, T+S)."""
    xP, yP = Pt
    if T[0] == S[0] and T[1] == S[1]:
        lam = fp2_mul(fp2_scale(fp2_sqr(T[0]), 3), fp2_inv(fp2_scale(T[1], 2)))
    else:
        lam = fp2_mul(fp2_sub(S[1], T[1]), fp2_inv(fp2_sub(S[0], T[0])))
    c = fp2_sub(fp2_mul(lam, T[0]), T[1])
    g = [(yP % P, 0), fp2_neg(fp2_scale(lam, xP)), FP2_ZERO, c, FP2_ZERO, FP2_ZERO]
    x3 = fp2_sub(fp2_sub(fp2_sqr(lam), T[0]), S[0])
    y3 = fp2_sub(fp2_mul(lam, fp2_sub(T[0], x3)), T[1])
    return fp12_from_w_basis(g), (x3, y3)


def miller_loop_single(Pt, Q):
    """Textbook optimal-ate Miller function for one pair (binary expansion of 6x+2)."""
    if Pt is None or Q is None:
        return FP12_ONE
    f = FP12_ONE
    T = Q
    for bit in bin(ATE_LOOP)[3:]:
        l, T2 = _line(T, T, Pt)
        f = fp12_mul(fp12_sqr(f), l)
        T = T2
        if bit == "1":
            l, T = _line(T, Q, Pt)
            f = fp12_mul(f, l)
    Q1 = g2_frobenius(Q, 1)
    Q2 = g2_neg(g2_frobenius(Q, 2))
    l, T = _line(T, Q1, Pt)
    f = fp12_mul(f, l)
    l, _ = _line(T, Q2, Pt)
    f = fp12_mul(f, l)
    return f


def miller_loop(Ps, Qs):
    """bn254.MillerLoop semantics: error on empty/mismatched input, pairs containing
    infinity skipped.  The raw value is only defined up to subfield factors."""
    if len(Ps) == 0 or len(Ps) != len(Qs):
        raise ValueError("invalid inputs sizes")
    f = FP12_ONE
    for Pt, Q in zip(Ps, Qs):
        f = fp12_mul(f, miller_loop_single(Pt, Q))
    return f


def final_exponentiation_literal(f):
    """f^(s*(p^12-1)/r) by literal exponentiation (2980-bit exponent)."""
    return fp12_pow(f, FE_EXPONENT)


def final_exponentiation(f):
    """Same value, computed easy-part-first (conj/inverse/Frobenius) then a literal pow of
    the hard exponent s*(p^4-p^2+1)/r.  ~4x faster than the literal form; the two are
    cross-checked in tests/test_oracle_ref.py."""
    t = fp12_mul(fp12_conj(f), fp12_inv(f))  # f^(p^6-1)
    t = fp12_mul(fp12_frobenius(t, 2), t)  # ^(p^2+1)
    return fp12_pow(t, FE_COFACTOR * ((P**4 - P * P + 1) // R))


def pair(Ps, Qs):
    """bn254.Pair(P[], Q[]) -> GT : product of pairings, one final exponentiation."""
    return final_exponentiation(miller_loop(Ps, Qs))


def pairing_check(Ps, Qs):
    return pair(Ps, Qs) == FP12_ONE


def gt_exp(x, k):
    """GT.Exp(x, k): k == 0 -> 1, k < 0 -> inverse first; generic Fp12 (no subgroup assumption)."""
    if k == 0:
        return FP12_ONE
    if k < 0:
        return fp12_pow(fp12_inv(x), -k)
    return fp12_pow(x, k)


def gt_div(a, b):
    return fp12_mul(a, fp12_inv(b))


# ----------------------------------------------------------------------------- gnark memory layout
def fp_to_mont_bytes(a):
    """fp.Element / fr.Element memory image: 4 little-endian u64 limbs, Montgomery R=2^256."""
    return (a * MONT_R % P).to_bytes(32, "little")


def fp_from_mont_bytes(b):
    return int.from_bytes(b, "little") * pow(MONT_R, -1, P) % P


def g1_to_bytes(pt):
    if pt is None:
        return bytes(64)
    return fp_to_mont_bytes(pt[0]) + fp_to_mont_bytes(pt[1])


def g1_from_bytes(b):
    if b == bytes(64):
        return None
    return (fp_from_mont_bytes(b[:32]), fp_from_mont_bytes(b[32:64]))


def g2_to_bytes(pt):
    if pt is None:
        return bytes(128)
    (x0, x1), (y0, y1) = pt
    return b"".join(fp_to_mont_bytes(c) for c in (x0, x1, y0, y1))


def g2_from_bytes(b):
    if b == bytes(128):
        return None
    c = [fp_from_mont_bytes(b[i * 32 : i * 32 + 32]) for i in range(4)]
    return ((c[0], c[1]), (c[2], c[3]))


def gt_to_bytes(a):
    return b"".join(fp_to_mont_bytes(c) for c in fp12_coeffs(a))


def gt_from_bytes(b):
    return fp12_from_coeffs([fp_from_mont_bytes(b[i * 32 : i * 32 + 32]) for i in range(12)])


def scalar_to_bytes(k):
    """C-ABI scalar: 32-byte little-endian unsigned integer (regular form, not Montgomery)."""
    return int(k).to_bytes(32, "little")


# ----------------------------------------------------------------------------- deterministic inputs
class SplitMix64:
    """Seeded generator for synthetic inputs."""

    def __init__(self, seed):
        self.s = seed & 0xFFFFFFFFFFFFFFFF

    def next(self):
        self.s = (self.s + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return z ^ (z >> 31)

    def u256(self):
        return sum(self.next() << (64 * i) for i in range(4))

    def scalar(self):
        return self.u256() % R

    def fp(self):
        return self.u256() % P


EDGE_SCALARS = [0, 1, 2, R - 1, R - 2, 1 << 128, LAMBDA_GLV]
