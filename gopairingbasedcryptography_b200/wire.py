"""Byte encodings of the gnark types: host-side conversions between the in-memory
Montgomery layout the engine works on and gnark's wire formats, needed by hash.FromGT (hash/hash_from_gt.go:5-8),
the Gentry06 transcript hash (ibe/gentry06_ibe/gentry06_ibe.go:319-343) and serialization/serialization_curve.go:5-33.

UNPINNED like the rest of the oracle story: gnark cannot run here, so the layouts follow the published
gnark-crypto conventions as recorded in SURVEY.md:
  fp.Element.Bytes()          32 B big-endian, regular (non-Montgomery) value
  GT.Bytes() / Marshal()      12 x 32 B big-endian regular coefficients, order C1.B2.A1, C1.B2.A0, C1.B1.A1, ... , C0.B0.A0
  G1Affine.Marshal()          64 B uncompressed  X || Y ; infinity = 64 zero bytes (BN254 has two spare bits: the only
                              infinity flag, 0b01, belongs to the COMPRESSED form; RawBytes writes mUncompressed = 0)
  G1Affine.Bytes()            32 B compressed X with the two top bits: 0b10 y "smallest", 0b11 y "largest", 0b01 infinity
  G2Affine.Marshal()         128 B uncompressed  X.A1 || X.A0 || Y.A1 || Y.A0
  G2Affine.Bytes()            64 B compressed  X.A1 || X.A0 with the same flag bits
Pure integer work on the host; nothing here touches the GPU."""
from __future__ import annotations

P = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
_RINV = pow(1 << 256, -1, P)
_RINV_R = pow(1 << 256, -1, R)
M_UNCOMPRESSED, M_INFINITY, M_SMALLEST, M_LARGEST, M_MASK = 0x00, 0x40, 0x80, 0xC0, 0xC0


def fp_from_mont(raw32):
    return int.from_bytes(raw32, "little") * _RINV % P


def fp_to_mont(v):
    return ((v % P) << 256) % P


def fp_mont_raw(v):
    return fp_to_mont(v).to_bytes(32, "little")


def fr_bytes(raw32):
    """fr.Element.Bytes(): 32 B big-endian regular form from the Montgomery memory image."""
    return (int.from_bytes(raw32, "little") * _RINV_R % R).to_bytes(32, "big")


def fr_set_bytes(b):
    """fr.Element.SetBytes: big-endian integer of any length reduced mod r -> Montgomery memory image."""
    return (((int.from_bytes(b, "big") % R) << 256) % R).to_bytes(32, "little")


def gt_bytes(raw384):
    c = [fp_from_mont(raw384[32 * i:32 * i + 32]) for i in range(12)]  # memory order C0.B0.A0, C0.B0.A1, C0.B1.A0 ...
    return b"".join(c[i].to_bytes(32, "big") for i in reversed(range(12)))


def gt_from_bytes(b384):
    c = [int.from_bytes(b384[32 * i:32 * i + 32], "big") for i in range(12)]
    if any(v >= P for v in c):
        raise ValueError("invalid fp.Element encoding")
    return b"".join(fp_mont_raw(v) for v in reversed(c))


def _lex_largest_fp(y):
    return y > (P - 1) // 2


def _lex_largest_fp2(y0, y1):
    return _lex_largest_fp(y1) if y1 != 0 else _lex_largest_fp(y0)


def g1_marshal(raw64):
    if raw64 == bytes(64):
        return bytes(64)  # uncompressed infinity: flag bits 00 and zero coordinates
    x, y = fp_from_mont(raw64[:32]), fp_from_mont(raw64[32:])
    return x.to_bytes(32, "big") + y.to_bytes(32, "big")


def g1_bytes(raw64):
    if raw64 == bytes(64):
        return bytes([M_INFINITY]) + bytes(31)
    x, y = fp_from_mont(raw64[:32]), fp_from_mont(raw64[32:])
    out = bytearray(x.to_bytes(32, "big"))
    out[0] |= M_LARGEST if _lex_largest_fp(y) else M_SMALLEST
    return bytes(out)


def _fp_sqrt(a):
    r = pow(a, (P + 1) // 4, P)  # p = 3 mod 4
    return r if r * r % P == a % P else None


def g1_unmarshal(b):
    flag = b[0] & M_MASK
    if len(b) == 64 and flag == M_UNCOMPRESSED:
        x, y = int.from_bytes(b[:32], "big"), int.from_bytes(b[32:], "big")
        if x >= P or y >= P:
            raise ValueError("invalid fp.Element encoding")
        if (x, y) != (0, 0) and (y * y - x * x * x - 3) % P:
            raise ValueError("invalid point: subgroup check failed")  # G1 has cofactor 1: on the curve == in the subgroup
        return fp_mont_raw(x) + fp_mont_raw(y)
    if flag == M_INFINITY:
        if len(b) != 32 or b[0] != M_INFINITY or any(b[1:]):
            raise ValueError("invalid infinity point encoding")
        return bytes(64)
    if len(b) != 32:
        raise ValueError("invalid point encoding")
    x = int.from_bytes(bytes([b[0] & ~M_MASK & 0xFF]) + b[1:], "big")
    y = _fp_sqrt((x * x * x + 3) % P)
    if x >= P or y is None:
        raise ValueError("invalid compressed coordinate: square root doesn't exist")
    if _lex_largest_fp(y) != (flag == M_LARGEST):
        y = P - y
    return fp_mont_raw(x) + fp_mont_raw(y)


def g2_marshal(raw128):
    if raw128 == bytes(128):
        return bytes(128)
    c = [fp_from_mont(raw128[32 * i:32 * i + 32]) for i in range(4)]  # X.A0, X.A1, Y.A0, Y.A1
    return b"".join(v.to_bytes(32, "big") for v in (c[1], c[0], c[3], c[2]))


def g2_bytes(raw128):
    if raw128 == bytes(128):
        return bytes([M_INFINITY]) + bytes(63)
    c = [fp_from_mont(raw128[32 * i:32 * i + 32]) for i in range(4)]
    out = bytearray(c[1].to_bytes(32, "big") + c[0].to_bytes(32, "big"))
    out[0] |= M_LARGEST if _lex_largest_fp2(c[2], c[3]) else M_SMALLEST
    return bytes(out)


def _fp2_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def _fp2_sqrt(a):
    """Square root in Fp[u]/(u^2+1) by the norm method; None when a is not a square."""
    a0, a1 = a
    if a1 == 0:
        r = _fp_sqrt(a0)
        if r is not None:
            return (r, 0)
        r = _fp_sqrt(-a0 % P)
        return None if r is None else (0, r)
    n = _fp_sqrt((a0 * a0 + a1 * a1) % P)
    if n is None:
        return None
    inv2 = pow(2, -1, P)
    for s in (n, -n % P):
        t = (a0 + s) * inv2 % P
        x0 = _fp_sqrt(t)
        if x0 is not None and x0 != 0:
            x1 = a1 * pow(2 * x0, -1, P) % P
            if _fp2_mul((x0, x1), (x0, x1)) == (a0 % P, a1 % P):
                return (x0, x1)
    return None


_B2 = _fp2_mul((3, 0), (9 * pow(82, -1, P) % P, -pow(82, -1, P) % P))  # 3 / (9 + u)


def _g2_in_subgroup(x, y):
    """On the twist y^2 = x^3 + 3/(9+u) and of order r: [r](x, y) = infinity by Jacobian double-and-add over Fp2
    (gnark's G2Affine.IsInSubGroup answers the same question with the psi endomorphism)."""
    sq = lambda a: _fp2_mul(a, a)
    sub = lambda a, b: ((a[0] - b[0]) % P, (a[1] - b[1]) % P)
    add = lambda a, b: ((a[0] + b[0]) % P, (a[1] + b[1]) % P)
    x3 = _fp2_mul(sq(x), x)
    if sub(sq(y), add(x3, _B2)) != (0, 0):
        return False
    X, Y, Z = (0, 0), (1, 0), (0, 0)  # accumulator = infinity
    for bit in bin(R)[2:]:
        if Z != (0, 0):  # dbl-2009-l
            A, Bq = sq(X), sq(Y)
            C = sq(Bq)
            D = sub(sub(sq(add(X, Bq)), A), C)
            D = add(D, D)
            E = add(add(A, A), A)
            F = sq(E)
            Z3 = _fp2_mul(Y, Z)
            Z3 = add(Z3, Z3)
            X3 = sub(F, add(D, D))
            C8 = add(C, C); C8 = add(C8, C8); C8 = add(C8, C8)
            Y3 = sub(_fp2_mul(E, sub(D, X3)), C8)
            X, Y, Z = X3, Y3, Z3
        if bit == "1":
            if Z == (0, 0):
                X, Y, Z = x, y, (1, 0)
            else:  # mixed addition
                Z2 = sq(Z)
                U2 = _fp2_mul(x, Z2)
                S2 = _fp2_mul(_fp2_mul(y, Z2), Z)
                H, Rr = sub(U2, X), sub(S2, Y)
                if H == (0, 0):
                    if Rr == (0, 0):
                        return False  # would need a doubling: cannot happen for a point of order r before the last bit
                    X, Y, Z = (0, 0), (1, 0), (0, 0)
                    continue
                H2 = sq(H)
                H3 = _fp2_mul(H2, H)
                V = _fp2_mul(X, H2)
                X3 = sub(sub(sq(Rr), H3), add(V, V))
                Y3 = sub(_fp2_mul(Rr, sub(V, X3)), _fp2_mul(Y, H3))
                X, Y, Z = X3, Y3, _fp2_mul(Z, H)
    return Z == (0, 0)


def g2_unmarshal(b):
    flag = b[0] & M_MASK
    if len(b) == 128 and flag == M_UNCOMPRESSED:
        v = [int.from_bytes(b[32 * i:32 * i + 32], "big") for i in range(4)]  # X.A1, X.A0, Y.A1, Y.A0
        if any(t >= P for t in v):
            raise ValueError("invalid fp.Element encoding")
        if any(v) and not _g2_in_subgroup((v[1], v[0]), (v[3], v[2])):
            raise ValueError("invalid point: subgroup check failed")
        return b"".join(fp_mont_raw(t) for t in (v[1], v[0], v[3], v[2]))
    if flag == M_INFINITY:
        if len(b) != 64 or b[0] != M_INFINITY or any(b[1:]):
            raise ValueError("invalid infinity point encoding")
        return bytes(128)
    if len(b) != 64:
        raise ValueError("invalid point encoding")
    x1 = int.from_bytes(bytes([b[0] & ~M_MASK & 0xFF]) + b[1:32], "big")
    x0 = int.from_bytes(b[32:], "big")
    x = (x0, x1)
    rhs = _fp2_mul(_fp2_mul(x, x), x)
    rhs = ((rhs[0] + _B2[0]) % P, (rhs[1] + _B2[1]) % P)
    y = _fp2_sqrt(rhs)
    if x0 >= P or x1 >= P or y is None:
        raise ValueError("invalid compressed coordinate: square root doesn't exist")
    if _lex_largest_fp2(*y) != (flag == M_LARGEST):
        y = (-y[0] % P, -y[1] % P)
    if not _g2_in_subgroup(x, y):
        raise ValueError("invalid point: subgroup check failed")
    return b"".join(fp_mont_raw(t) for t in (x0, x1, y[0], y[1]))


# ---- the reference's consumers of these encodings ------------------------------------------------------------------
def hash_from_gt(raw384):
    """hash.FromGT (hash/hash_from_gt.go:5-8): the 384 bytes of GT.Bytes()."""
    return gt_bytes(raw384)


def gentry06_h(u_raw64, v_raw384, w_raw384):
    """Gentry06 transcript hash H: G1 x GT x GT -> Zp (ibe/gentry06_ibe/gentry06_ibe.go:319-343):
    beta = fr.SetBytes(SHA-256(u.Bytes() || v.Bytes() || w.Bytes())).  Returns the fr.Element memory image (32 B)."""
    import hashlib

    return fr_set_bytes(hashlib.sha256(g1_bytes(u_raw64) + gt_bytes(v_raw384) + gt_bytes(w_raw384)).digest())
