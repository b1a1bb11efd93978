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
