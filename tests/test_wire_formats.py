"""Wire formats: host-side byte encodings of the gnark types.  Unpinned against gnark
(cannot run here); checked for round trips, flag bits and the layout rules SURVEY.md records."""
import pytest

from gopairingbasedcryptography_b200 import bn254, wire
from oracle import bn254_ref as o


def test_fp_and_gt_bytes_layout():
    e = o.pair([o.G1_GEN], [o.G2_GEN])
    raw = o.gt_to_bytes(e)
    b = wire.gt_bytes(raw)
    assert len(b) == 384
    # first 32 bytes = C1.B2.A1, last 32 = C0.B0.A0, big-endian regular form
    assert int.from_bytes(b[:32], "big") == e[1][2][1] and int.from_bytes(b[-32:], "big") == e[0][0][0]
    assert wire.gt_from_bytes(b) == raw
    assert bn254.GT(raw).Bytes() == b and bn254.GT().Unmarshal(b).raw == raw
    with pytest.raises(ValueError):
        wire.gt_from_bytes(b"\xff" * 384)


def test_fr_bytes():
    assert wire.fr_bytes(wire.fr_set_bytes((5).to_bytes(32, "big"))) == (5).to_bytes(32, "big")
    # SetBytes reduces arbitrary-length big-endian input mod r (hash/hash_to.go:30-35 relies on it)
    big = (o.R + 7).to_bytes(40, "big")
    assert int.from_bytes(wire.fr_bytes(wire.fr_set_bytes(big)), "big") == 7


def test_g1_encodings_round_trip():
    for k in (1, 2, 12345, o.R - 1):
        pt = o.g1_mul(o.G1_GEN, k)
        raw = o.g1_to_bytes(pt)
        m, c = wire.g1_marshal(raw), wire.g1_bytes(raw)
        assert len(m) == 64 and len(c) == 32
        assert int.from_bytes(m[:32], "big") == pt[0] and int.from_bytes(m[32:], "big") == pt[1]
        assert (c[0] & 0xC0) in (0x80, 0xC0) and ((c[0] & 0xC0) == 0xC0) == (pt[1] > (o.P - 1) // 2)
        assert wire.g1_unmarshal(m) == raw and wire.g1_unmarshal(c) == raw
        assert bn254.G1Affine().Unmarshal(bn254.G1Affine(raw).Bytes()).raw == raw
    inf = bytes(64)
    assert wire.g1_bytes(inf)[0] == 0x40 and wire.g1_unmarshal(wire.g1_bytes(inf)) == inf
    assert wire.g1_unmarshal(wire.g1_marshal(inf)) == inf
    # -P flips the sign flag only
    p1, p2 = o.g1_mul(o.G1_GEN, 77), o.g1_neg(o.g1_mul(o.G1_GEN, 77))
    c1, c2 = wire.g1_bytes(o.g1_to_bytes(p1)), wire.g1_bytes(o.g1_to_bytes(p2))
    assert c1[1:] == c2[1:] and (c1[0] ^ c2[0]) == 0x40


def test_g2_encodings_round_trip():
    for k in (1, 3, 987654321, o.R - 2):
        pt = o.g2_mul(o.G2_GEN, k)
        raw = o.g2_to_bytes(pt)
        m, c = wire.g2_marshal(raw), wire.g2_bytes(raw)
        assert len(m) == 128 and len(c) == 64
        # uncompressed order X.A1 || X.A0 || Y.A1 || Y.A0
        assert [int.from_bytes(m[32 * i:32 * i + 32], "big") for i in range(4)] == [pt[0][1], pt[0][0], pt[1][1], pt[1][0]]
        assert wire.g2_unmarshal(m) == raw and wire.g2_unmarshal(c) == raw
        assert bn254.G2Affine().Unmarshal(bn254.G2Affine(raw).Bytes()).raw == raw
    inf = bytes(128)
    assert wire.g2_unmarshal(wire.g2_bytes(inf)) == inf and wire.g2_unmarshal(wire.g2_marshal(inf)) == inf
    with pytest.raises(ValueError):
        wire.g2_unmarshal(bytes([0x80]) + bytes(62) + b"\x05")  # x with no point on the twist (or non-square)


def test_uncompressed_infinity_is_all_zero_and_validation():
    """gnark RawBytes() of the point at infinity on BN254 writes mUncompressed (0b00) and zero coordinates -- 0b01 is
    the COMPRESSED infinity flag.  Unmarshal runs gnark's curve / subgroup checks."""
    assert wire.g1_marshal(bytes(64)) == bytes(64) and wire.g2_marshal(bytes(128)) == bytes(128)
    assert wire.g1_bytes(bytes(64)) == bytes([0x40]) + bytes(31) and wire.g2_bytes(bytes(128)) == bytes([0x40]) + bytes(63)
    with pytest.raises(ValueError):  # (1, 3) is not on y^2 = x^3 + 3
        wire.g1_unmarshal((1).to_bytes(32, "big") + (3).to_bytes(32, "big"))
    with pytest.raises(ValueError):  # infinity flag with trailing garbage
        wire.g1_unmarshal(bytes([0x40]) + bytes(30) + b"\x01")
    # a point ON the twist but OUTSIDE the order-r subgroup (the twist has cofactor 2p - r): find one by x-increment
    x = (1, 0)
    while True:
        rhs = wire._fp2_mul(wire._fp2_mul(x, x), x)
        rhs = ((rhs[0] + wire._B2[0]) % o.P, (rhs[1] + wire._B2[1]) % o.P)
        y = wire._fp2_sqrt(rhs)
        if y is not None:
            break
        x = (x[0] + 1, 0)
    assert not wire._g2_in_subgroup(x, y)
    enc = b"".join(v.to_bytes(32, "big") for v in (x[1], x[0], y[1], y[0]))
    with pytest.raises(ValueError, match="subgroup"):
        wire.g2_unmarshal(enc)
    g = o.g2_mul(o.G2_GEN, 5)
    assert wire._g2_in_subgroup(g[0], g[1])


def test_reference_consumers_of_the_encodings():
    """hash.FromGT (hash/hash_from_gt.go:5-8) and Gentry06's h (ibe/gentry06_ibe/gentry06_ibe.go:319-343), restated
    independently here from the oracle's integer tuples."""
    import hashlib

    u = o.g1_mul(o.G1_GEN, 4242)
    v = o.pair([o.G1_GEN], [o.G2_GEN])
    w = o.pair([u], [o.G2_GEN])
    flat = lambda e: [e[c][b][a] for c in (1, 0) for b in (2, 1, 0) for a in (1, 0)]  # C1.B2.A1 ... C0.B0.A0
    vb = b"".join(t.to_bytes(32, "big") for t in flat(v))
    wb = b"".join(t.to_bytes(32, "big") for t in flat(w))
    assert wire.hash_from_gt(o.gt_to_bytes(v)) == vb
    ub = bytearray(u[0].to_bytes(32, "big"))
    ub[0] |= 0xC0 if u[1] > (o.P - 1) // 2 else 0x80
    beta = int.from_bytes(hashlib.sha256(bytes(ub) + vb + wb).digest(), "big") % o.R
    got = wire.gentry06_h(o.g1_to_bytes(u), o.gt_to_bytes(v), o.gt_to_bytes(w))
    assert int.from_bytes(got, "little") == beta * (1 << 256) % o.R
