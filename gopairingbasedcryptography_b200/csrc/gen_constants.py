"""Generate bn254_constants.cuh (8x32-bit-limb Montgomery constants for the CUDA engine).

Run from the repo root:  python gopairingbasedcryptography_b200/csrc/gen_constants.py
Self-contained on purpose: the product never imports oracle/.  Everything is derived from the
curve seed x0 and xi = 9+u with plain Python integers.
"""
from __future__ import annotations

import os

X0 = 4965661367192848881
P = 36 * X0**4 + 36 * X0**3 + 24 * X0**2 + 6 * X0 + 1
R = 36 * X0**4 + 36 * X0**3 + 18 * X0**2 + 6 * X0 + 1
MONT = 1 << 256


def f2mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def f2pow(a, e):
    out = (1, 0)
    while e:
        if e & 1:
            out = f2mul(out, a)
        a = f2mul(a, a)
        e >>= 1
    return out


def f2inv(a):
    n = pow(a[0] * a[0] + a[1] * a[1], -1, P)
    return (a[0] * n % P, -a[1] * n % P)


def limbs32(v, n=8):
    return ", ".join("0x%08xu" % ((v >> (32 * i)) & 0xFFFFFFFF) for i in range(n))


def m(v):
    return v * MONT % P


def fp_init(v):
    return "{{%s}}" % limbs32(m(v))


def fp2_init(a):
    return "{%s, %s}" % (fp_init(a[0]), fp_init(a[1]))


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


def main():
    xi = (9, 1)
    b2 = f2mul((3, 0), f2inv(xi))
    g1 = [f2pow(xi, j * (P - 1) // 6) for j in range(6)]
    g2 = [f2pow(xi, j * (P * P - 1) // 6) for j in range(6)]
    g3 = [f2pow(xi, j * (P**3 - 1) // 6) for j in range(6)]
    assert all(g[1] == 0 for g in g2)
    G2X = (10857046999023057135944570762232829481370756359578518086990519993285655852781,
           11559732032986387107991004021392285783925812861821192530917403151452391805634)
    G2Y = (8495653923123431417604973247489272438418190587263600148770280649306958101930,
           4082367875863433681332203403145435568316851327593401208105741076214120093531)
    lam = 36 * X0**3 + 18 * X0**2 + 6 * X0 + 1
    assert (lam * lam + lam + 1) % R == 0
    # beta: primitive cube root of unity in Fp with phi(x,y) = (beta x, y) = [lam](x,y) on G1;
    # fixed by checking on the generator (1,2) in tests (tests/test_constants.py).
    beta_candidates = [pow(c, (P - 1) // 3, P) for c in (2, 3, 5, 7) if pow(c, (P - 1) // 3, P) != 1]
    beta = beta_candidates[0]
    # ---- GLV: pick beta with phi(x,y) = (beta x, y) = [lam](x,y) on G1 and on the twist, by testing it
    def ec_add(a, b, f_mul, f_sub, f_inv, f_zero):
        if a is None:
            return b
        if b is None:
            return a
        if a[0] == b[0]:
            if f_sub(a[1], f_sub(f_zero, b[1])) == f_zero:
                return None
            three_x2 = f_mul(f_mul(a[0], a[0]), THREE[type(f_zero)])
            lam_ = f_mul(three_x2, f_inv(f_mul(a[1], TWO[type(f_zero)])))
        else:
            lam_ = f_mul(f_sub(b[1], a[1]), f_inv(f_sub(b[0], a[0])))
        x3 = f_sub(f_sub(f_mul(lam_, lam_), a[0]), b[0])
        return (x3, f_sub(f_mul(lam_, f_sub(a[0], x3)), a[1]))

    def ec_mul(pt, k, ops):
        acc = None
        while k:
            if k & 1:
                acc = ec_add(acc, pt, *ops)
            pt = ec_add(pt, pt, *ops)
            k >>= 1
        return acc

    THREE = {int: 3, tuple: (3, 0)}
    TWO = {int: 2, tuple: (2, 0)}
    fp_ops = (lambda a, b: a * b % P, lambda a, b: (a - b) % P, lambda a: pow(a, -1, P), 0)
    fp2_ops = (f2mul, lambda a, b: ((a[0] - b[0]) % P, (a[1] - b[1]) % P), f2inv, (0, 0))
    g1_lam = ec_mul((1, 2), lam, fp_ops)
    beta_g1 = [b for b in (beta, beta * beta % P) if (b * 1 % P, 2) == g1_lam]
    assert len(beta_g1) == 1
    beta_g1 = beta_g1[0]
    g2_lam = ec_mul((G2X, G2Y), lam, fp2_ops)
    beta_g2 = [b for b in (beta, beta * beta % P) if ((G2X[0] * b % P, G2X[1] * b % P), G2Y) == g2_lam]
    assert len(beta_g2) == 1
    beta_g2 = beta_g2[0]
    # short lattice basis of {(a,b): a + b*lam = 0 mod r} by the extended Euclidean algorithm
    rows = [(R, 1, 0), (lam, 0, 1)]
    while rows[-1][0] != 0:
        q = rows[-2][0] // rows[-1][0]
        rows.append(tuple(x - q * y for x, y in zip(rows[-2], rows[-1])))
    import math
    sq = math.isqrt(R)
    l = max(i for i, row in enumerate(rows) if row[0] >= sq)
    v1 = (rows[l + 1][0], -rows[l + 1][2])
    cand = [(rows[l][0], -rows[l][2]), (rows[l + 2][0], -rows[l + 2][2])]
    v2 = min(cand, key=lambda v: v[0] * v[0] + v[1] * v[1])
    for a_, b_ in (v1, v2):
        assert (a_ + b_ * lam) % R == 0
    (a1, lb1), (a2, lb2) = v1, v2
    det = a1 * lb2 - a2 * lb1
    assert abs(det) == R
    # k = c1 v1 + c2 v2 + (k1,k2):  c1 = round(k b2/det), c2 = round(-k b1/det); device uses floor(k*g/2^256)
    sgn = 1 if det > 0 else -1
    n1, n2 = sgn * lb2, -sgn * lb1          # c1 ~ k*n1/r, c2 ~ k*n2/r
    g1c, g2c = (abs(n1) << 256) // R, (abs(n2) << 256) // R
    M256 = (1 << 256) - 1
    s1, s2 = (1 if n1 >= 0 else -1), (1 if n2 >= 0 else -1)
    # k1 = k - c1 a1 - c2 a2 ; k2 = -c1 b1 - c2 b2, with c_i = s_i * floor(k*g_ic >> 256)
    glv = {"G1C": g1c, "G2C": g2c, "K1_M1": (-s1 * a1) & M256, "K1_M2": (-s2 * a2) & M256,
           "K2_M1": (-s1 * lb1) & M256, "K2_M2": (-s2 * lb2) & M256}
    # self-check of the exact device formulas on many scalars
    import random
    rnd = random.Random(1)
    worst = 0
    for k in [0, 1, 2, R - 1, R - 2, R, R + 1, 1 << 128, lam, (1 << 256) - 1] + [rnd.randrange(1 << 256) for _ in range(20000)]:
        c1, c2 = (k * g1c) >> 256, (k * g2c) >> 256
        k1 = (k + c1 * glv["K1_M1"] + c2 * glv["K1_M2"]) & M256
        k2 = (c1 * glv["K2_M1"] + c2 * glv["K2_M2"]) & M256
        k1 = k1 - (1 << 256) if k1 >> 255 else k1
        k2 = k2 - (1 << 256) if k2 >> 255 else k2
        assert (k1 + k2 * lam - k) % R == 0
        worst = max(worst, abs(k1).bit_length(), abs(k2).bit_length())
    assert worst <= 131, worst
    # ---- 4-dimensional GLS decomposition on G2 (Galbraith-Scott): psi = twist . Frobenius . untwist acts on G2 as
    # multiplication by lam4 = p mod r = 6 x0^2, a root of X^4 - X^2 + 1; k = k0 + k1 lam4 + k2 lam4^2 + k3 lam4^3 with
    # |k_j| < 2^67 by Babai rounding against the basis below (rows are lattice vectors: sum row_j lam4^j = 0 mod r;
    # their determinant is -3r -- an index-3 sublattice, which costs nothing but a slightly larger bound)
    lam4 = P % R
    assert lam4 == 6 * X0 * X0 and (lam4**4 - lam4**2 + 1) % R == 0
    psi_pt = ((G2X[0], -G2X[1] % P), (G2Y[0], -G2Y[1] % P))          # conjugates
    psi_pt = (f2mul(psi_pt[0], g1[2]), f2mul(psi_pt[1], g1[3]))      # psi(x, y) = (conj(x) xi^((p-1)/3), conj(y) xi^((p-1)/2))
    assert psi_pt == ec_mul((G2X, G2Y), lam4, fp2_ops), "psi is not multiplication by p mod r on G2"
    x = X0
    B4 = [[x + 1, x, x, -2 * x], [2 * x + 1, -x, -(x + 1), -x], [2 * x, 2 * x + 1, 2 * x + 1, 2 * x + 1], [x - 1, 4 * x + 2, -(2 * x - 1), x - 1]]
    for row in B4:
        assert sum(c * lam4**j for j, c in enumerate(row)) % R == 0
    from fractions import Fraction

    def mat_inv(M):
        n = len(M)
        A = [[Fraction(v) for v in r] + [Fraction(int(i == j)) for j in range(n)] for i, r in enumerate(M)]
        for i in range(n):
            pv = next(r for r in range(i, n) if A[r][i] != 0)
            A[i], A[pv] = A[pv], A[i]
            d = A[i][i]
            A[i] = [v / d for v in A[i]]
            for r in range(n):
                if r != i and A[r][i] != 0:
                    f = A[r][i]
                    A[r] = [a - f * b for a, b in zip(A[r], A[i])]
        return [r[n:] for r in A]

    row0 = mat_inv(B4)[0]                       # (k, 0, 0, 0) B^-1 = k row0, row0_i = alpha_i / (3r)
    alphas = [int(r_ * 3 * R) for r_ in row0]
    assert all(Fraction(a, 3 * R) == r_ for a, r_ in zip(alphas, row0))
    g4 = [(abs(a) << 256) // (3 * R) for a in alphas]
    s4 = [1 if a >= 0 else -1 for a in alphas]
    # k_j = [j == 0] k + sum_i c_i M4[i][j] (mod 2^256, two's complement), c_i = floor(k g4_i / 2^256)
    M4 = [[(-s4[i] * B4[i][j]) & M256 for j in range(4)] for i in range(4)]
    worst4 = 0
    for k in [0, 1, 2, R - 1, R - 2, R, R + 1, 1 << 128, lam4, (1 << 256) - 1] + [rnd.randrange(1 << 256) for _ in range(20000)]:
        c = [(k * gi) >> 256 for gi in g4]
        kv = [((k if j == 0 else 0) + sum(c[i] * M4[i][j] for i in range(4))) & M256 for j in range(4)]
        kv = [v - (1 << 256) if v >> 255 else v for v in kv]
        assert sum(kj * lam4**j for j, kj in enumerate(kv)) % R == k % R
        worst4 = max(worst4, max(abs(v).bit_length() for v in kv))
    assert worst4 <= 67, worst4
    ate_naf = naf(6 * X0 + 2)
    x0_naf3 = naf(X0, 3)

    o = []
    w = o.append
    w("// GENERATED by gen_constants.py -- do not edit.")
    w("#pragma once")
    w("namespace bn254 {")
    for i in range(8):
        w("static constexpr uint32_t P%d = 0x%08xu;" % (i, (P >> (32 * i)) & 0xFFFFFFFF))
    w("static constexpr uint32_t P_INV32 = 0x%08xu;  // -p^-1 mod 2^32" % ((-pow(P, -1, 1 << 32)) % (1 << 32)))
    w("#define BN254_FP_ONE %s" % fp_init(1))
    w("#define BN254_FP_R3 %s  // R^3 mod p as plain limbs: (x R)^-1 * R^3 / R = x^-1 R (binary inversion, wvm.cuh)" % fp_init(pow(2, 512, P)))
    w("#define BN254_FP_R2 {{%s}}" % limbs32(MONT * MONT % P))
    w("BN_CONST uint32_t FP_PM2[8] = {%s};  // p-2, inversion exponent" % limbs32(P - 2))
    for i in range(16):
        w("static constexpr uint32_t PSQ%d = 0x%08xu;" % (i, ((P * P) >> (32 * i)) & 0xFFFFFFFF))
    w("BN_CONST uint32_t FR_MOD[8] = {%s};" % limbs32(R))
    w("BN_CONST Fp2 TWIST_3B = %s;  // 3*3/(9+u)" % fp2_init(((3 * b2[0]) % P, (3 * b2[1]) % P)))
    w("BN_CONST Fp2 TWIST_B = %s;" % fp2_init(b2))
    for name, tab in (("GAMMA1", g1), ("GAMMA3", g3)):
        w("BN_CONST Fp2 %s[6] = {" % name)
        for g in tab:
            w("  %s," % fp2_init(g))
        w("};")
    w("BN_CONST Fp GAMMA2[6] = {")
    for g in g2:
        w("  %s," % fp_init(g[0]))
    w("};")
    w("BN_CONST Fp G1_GEN_X = %s;" % fp_init(1))
    w("BN_CONST Fp G1_GEN_Y = %s;" % fp_init(2))
    w("BN_CONST Fp2 G2_GEN_X = %s;" % fp2_init(G2X))
    w("BN_CONST Fp2 G2_GEN_Y = %s;" % fp2_init(G2Y))
    w("BN_CONST Fp GLV_BETA = %s;  // phi(x,y) = (beta x, y) = [lambda](x,y) on G1" % fp_init(beta_g1))
    w("BN_CONST Fp GLV_BETA_G2 = %s;  // same on the twist (G2)" % fp_init(beta_g2))
    for name, v in glv.items():
        w("BN_CONST uint32_t GLV_%s[8] = {%s};" % (name, limbs32(v)))
    for i in range(4):
        w("BN_CONST uint32_t GLS4_G%d[8] = {%s};" % (i, limbs32(g4[i])))
        for j in range(4):
            w("BN_CONST uint32_t GLS4_M%d%d[8] = {%s};" % (i, j, limbs32(M4[i][j])))
    w("static constexpr int GLS4_MAX_BITS = %d;  // |k_j| < 2^GLS4_MAX_BITS for every 256-bit scalar" % (worst4 + 1))
    w("static constexpr int GLV_MAX_BITS = %d;  // |k1|, |k2| < 2^GLV_MAX_BITS for every 256-bit scalar" % (worst + 1))
    w("static constexpr int ATE_NAF_LEN = %d;" % len(ate_naf))
    w("BN_CONST signed char ATE_NAF[%d] = {%s};" % (len(ate_naf), ",".join(map(str, ate_naf))))
    w("static constexpr int X0_NAF3_LEN = %d;" % len(x0_naf3))
    w("BN_CONST signed char X0_NAF3[%d] = {%s};  // width-3 signed digits of x0, LSB first" % (len(x0_naf3), ",".join(map(str, x0_naf3))))
    w("static constexpr unsigned long long X0_SEED = %dull;" % X0)
    # ---- hash-to-curve (RFC 9380 SVDW; gnark hash_to_g1.go / hash_to_g2.go): constants derived from Z
    def fsqrt(a):
        r = pow(a % P, (P + 1) // 4, P)
        assert r * r % P == a % P
        return r

    def fissq(a):
        return a % P == 0 or pow(a % P, (P - 1) // 2, P) == 1

    def f2sqrt(a):
        if a[1] % P == 0:
            return (fsqrt(a[0]), 0) if fissq(a[0]) else (0, fsqrt(-a[0] % P))
        n = fsqrt((a[0] * a[0] + a[1] * a[1]) % P)
        x2 = (a[0] + n) * pow(2, -1, P) % P
        if not fissq(x2):
            x2 = (a[0] - n) * pow(2, -1, P) % P
        x = fsqrt(x2)
        r = (x, a[1] * pow(2 * x, -1, P) % P)
        assert f2mul(r, r) == (a[0] % P, a[1] % P)
        return r

    z1 = 1
    gz = (z1**3 + 3) % P
    c3 = fsqrt(-gz * 3 * z1 * z1 % P)
    c3 = P - c3 if c3 & 1 else c3
    h1 = {"Z": z1, "C1": gz, "C2": -z1 * pow(2, -1, P) % P, "C3": c3, "C4": -4 * gz * pow(3 * z1 * z1, -1, P) % P}
    assert h1["C3"] == 8815841940592487685674414971303048083897117035520822607866  # gnark bn254 HashE1 c3
    z2 = (0, 1)
    gz2 = f2mul(f2mul(z2, z2), z2)
    gz2 = ((gz2[0] + b2[0]) % P, (gz2[1] + b2[1]) % P)
    z2sq3 = tuple(3 * c % P for c in f2mul(z2, z2))
    c3 = f2sqrt(tuple(-c % P for c in f2mul(gz2, z2sq3)))
    sgn = (c3[0] & 1) | (int(c3[0] == 0) & (c3[1] & 1))
    c3 = tuple(-c % P for c in c3) if sgn else c3
    c4 = tuple(-c % P for c in f2mul(tuple(4 * c % P for c in gz2), f2inv(z2sq3)))
    h2 = {"Z": z2, "C1": gz2, "C2": tuple(-c * pow(2, -1, P) % P for c in z2), "C3": c3, "C4": c4}
    for k, v in h1.items():
        w("BN_CONST Fp H2C_G1_%s = %s;" % (k, fp_init(v)))
    for k, v in h2.items():
        w("BN_CONST Fp2 H2C_G2_%s = %s;" % (k, fp2_init(v)))
    w("BN_CONST Fp CURVE_B = %s;" % fp_init(3))
    w("BN_CONST uint32_t FP_PM1H[8] = {%s};  // (p-1)/2, Euler criterion" % limbs32((P - 1) // 2))
    w("BN_CONST uint32_t FP_PP1Q[8] = {%s};  // (p+1)/4, square root (p = 3 mod 4)" % limbs32((P + 1) // 4))
    w("BN_CONST uint32_t FP_PM3Q[8] = {%s};  // (p-3)/4: c^((p-3)/4) gives a square root (times c) and its inverse at once" % limbs32((P - 3) // 4))
    w("BN_CONST Fp FP_HALF = %s;" % fp_init(pow(2, -1, P)))
    # hash_to_field: a 48-byte big-endian integer c2 2^256 + c1 2^128 + c0 enters Montgomery form as
    # mont(c0, R^2) + mont(c1, 2^128 R^2) + mont(c2, 2^256 R^2)  (raw limbs, every c_i < 2^128 < p)
    for k in range(3):
        w("BN_CONST Fp H2F_K%d = {{%s}};" % (k, limbs32((1 << (128 * k)) * MONT * MONT % P)))
    w("BN_CONST Fp FP_RAW_ONE = {{%s}};  // mont(a, 1) = a/R: Montgomery -> regular form" % limbs32(1))
    w("}  // namespace bn254")
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bn254_constants.cuh")
    with open(path, "w") as f:
        f.write("\n".join(o) + "\n")
    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, "generators_g1.inc"), "w") as f:
        f.write(limbs32(m(1)) + ",\n" + limbs32(m(2)) + "\n")
    with open(os.path.join(here, "generators_g2.inc"), "w") as f:
        f.write(",\n".join(limbs32(m(c)) for c in (G2X[0], G2X[1], G2Y[0], G2Y[1])) + "\n")
    print("wrote", path, "x0 naf3 weight", sum(1 for d in x0_naf3 if d), "len", len(x0_naf3))


if __name__ == "__main__":
    main()
