"""Hash-to-curve (hash/hash_to.go -> gnark bn254.HashToG1 / HashToG2).

CPU part: the oracle against the RFC 9380 expand_message_xmd vectors, its algebraic properties, the committed golden
vectors, and the DEVICE code (csrc/hash_to_curve.cuh compiled for the host) against the oracle.  GPU part: the C ABI
against the oracle and a BLS sign / verify round trip with the real H(m) (signature/bls01_signature/bls_signature.go)."""
import ctypes
import json
import os

import numpy as np
import pytest

from oracle import bn254_ref as o
from oracle import hash_to_curve_ref as h
from oracle import port

import common
from test_emu_device_code import emu  # noqa: F401  (fixture)

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def vectors():
    with open(os.path.join(HERE, "golden", "hash_to_curve_vectors.json")) as f:
        return json.load(f)


def raw_mul_g2(pt, k):
    acc = None
    while k:
        if k & 1:
            acc = o.g2_add(acc, pt)
        pt = o.g2_add(pt, pt)
        k >>= 1
    return acc


def test_expand_message_xmd_rfc9380_vectors(vectors):
    v = vectors["rfc9380_k1_expand_message_xmd_sha256"]
    for e in v["vectors"]:
        assert h.expand_message_xmd(e["msg"].encode(), v["dst"].encode(), e["len"]).hex() == e["uniform_bytes"]
    # length / tag limits of the RFC (gnark returns an error)
    with pytest.raises(ValueError):
        h.expand_message_xmd(b"x", b"d" * 256, 32)


def test_oracle_properties():
    rng = o.SplitMix64(2024)
    for i in range(6):
        msg = bytes(rng.next() & 0xFF for _ in range(i * 7))
        p = h.hash_to_g1(msg, h.DST_BYTES_G1)
        assert o.g1_on_curve(p) and p == h.hash_to_g1(msg, h.DST_BYTES_G1)
        assert p != h.hash_to_g1(msg, h.DST_STRING_G1)  # domain separation
        q = h.hash_to_g2(msg, h.DST_BYTES_G2)
        assert o.g2_on_curve(q) and raw_mul_g2(q, o.R) is None  # order-r subgroup after cofactor clearing
    # the SVDW map alone lands on the twist but not in the subgroup; psi acts as multiplication by p on G2
    u = h.hash_to_fp(b"abc", h.DST_BYTES_G2, 4)
    q0 = h.map_to_curve_g2((u[0], u[1]))
    assert o.g2_on_curve(q0) and raw_mul_g2(q0, o.R) is not None
    assert h.g2_psi(o.G2_GEN) == o.g2_mul(o.G2_GEN, o.P % o.R)
    # sgn0 convention of the map: sgn0(y) == sgn0(u)
    for uu in (1, 2, 12345, o.P - 1):
        x, y = h.map_to_curve_g1(uu)
        assert (y * y - x**3 - 3) % o.P == 0 and (y & 1) == (uu & 1)


def test_oracle_matches_golden(vectors):
    for e in vectors["g1"]:
        assert o.g1_to_bytes(h.hash_to_g1(bytes.fromhex(e["msg"]), e["dst"].encode())).hex() == e["point"]
    for e in vectors["g2"][:4]:
        assert o.g2_to_bytes(h.hash_to_g2(bytes.fromhex(e["msg"]), e["dst"].encode())).hex() == e["point"]


def _emu(emu, fn, msg, dst, nbytes):
    out = (ctypes.c_uint8 * nbytes)()
    getattr(emu, fn)(msg, ctypes.c_size_t(len(msg)), dst, ctypes.c_size_t(len(dst)), out)
    return bytes(out)


def test_device_code_on_host_matches_golden(emu, vectors):  # noqa: F811
    for e in vectors["g1"]:
        assert _emu(emu, "emu_hash_to_g1", bytes.fromhex(e["msg"]), e["dst"].encode(), 64).hex() == e["point"]
    for e in vectors["g2"]:
        assert _emu(emu, "emu_hash_to_g2", bytes.fromhex(e["msg"]), e["dst"].encode(), 128).hex() == e["point"]
    # field elements, a 255-byte tag and an empty tag
    for dst in (b"", b"t" * 255):
        f = _emu(emu, "emu_hash_to_field", b"message", dst, 128)
        assert [o.fp_from_mont_bytes(f[32 * i:32 * i + 32]) for i in range(4)] == h.hash_to_fp(b"message", dst, 4)


def test_device_quadratic_character_is_the_jacobi_symbol(emu):  # noqa: F811
    """fp_is_square of the device code (binary Jacobi symbol, no field multiplication) against Euler's criterion on Python
    integers: edge values, small numbers, powers of two (long runs of trailing zeros, whole zero limbs) and 400 random ones."""
    rng = o.SplitMix64(0x1ACB1)
    vals = [0, 1, 2, 3, 4, 5, 7, 8, o.P - 1, o.P - 2, (o.P - 1) // 2, (o.P + 1) // 2, 1 << 32, 1 << 64, 1 << 200, 3 << 96, (1 << 253) + 1]
    vals += [rng.fp() for _ in range(400)]
    # the kernel sees Montgomery residues v R mod p; R is a square, so the character is that of v
    buf = np.frombuffer(b"".join(o.fp_to_mont_bytes(v) for v in vals), dtype=np.uint8).copy()
    out = (ctypes.c_int * len(vals))()
    emu.emu_fp_is_square(buf.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(len(vals)), out)
    want = [1 if v == 0 or pow(v, (o.P - 1) // 2, o.P) == 1 else 0 for v in vals]
    assert list(out) == want
    # and on the raw limbs themselves (the routine does not care which representation it is handed)
    raw = np.frombuffer(b"".join(int(v).to_bytes(32, "little") for v in vals), dtype=np.uint8).copy()
    emu.emu_fp_is_square(raw.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(len(vals)), out)
    assert list(out) == want


def test_device_g2_square_root_takes_both_branches(emu):  # noqa: F811
    """The Fp2 square root of the device code derives the root from ONE ladder c^((p-3)/4) whichever of its two candidates
    is the square (hash_to_curve.cuh fp2_sqrt): 40 more messages against the definitional oracle -- both branches of both
    SVDW maps are taken many times over (each is a coin flip per map)."""
    dst = h.DST_BYTES_G2
    for i in range(40):
        msg = b"square-root branch %d" % i
        assert _emu(emu, "emu_hash_to_g2", msg, dst, 128) == o.g2_to_bytes(h.hash_to_g2(msg, dst)), i


@pytest.mark.gpu
def test_gpu_hash_to_curve_vs_oracle(engine, vectors):
    for grp, fn in (("g1", engine.hash_to_g1_batch), ("g2", engine.hash_to_g2_batch)):
        by_dst = {}
        for e in vectors[grp]:
            by_dst.setdefault(e["dst"], []).append(e)
        for dst, es in by_dst.items():
            got = fn([bytes.fromhex(e["msg"]) for e in es], dst.encode())
            for g, e in zip(got, es):
                assert g.tobytes().hex() == e["point"]
    rng = o.SplitMix64(77)
    msgs = [bytes(rng.next() & 0xFF for _ in range((13 * i) % 97)) for i in range(150)]  # ragged lengths, > 1 CTA
    g1 = engine.hash_to_g1_batch(msgs, h.DST_BYTES_G1)
    g2 = engine.hash_to_g2_batch(msgs, h.DST_BYTES_G2)
    for i in (0, 1, 7, 64, 127, 128, 149):
        assert g1[i].tobytes() == o.g1_to_bytes(h.hash_to_g1(msgs[i], h.DST_BYTES_G1))
        assert g2[i].tobytes() == o.g2_to_bytes(h.hash_to_g2(msgs[i], h.DST_BYTES_G2))
    # SHA-256 padding boundaries of the b_0 block (64-byte Z_pad + msg + 3 + DST'), long messages, longest DST
    edge = [bytes((7 * i + j) & 0xFF for j in range(ln)) for i, ln in enumerate((54, 55, 56, 63, 64, 65, 119, 120, 1000, 5000))]
    for dst in (b"D", b"d" * 255):
        e1 = engine.hash_to_g1_batch(edge, dst)
        e2 = engine.hash_to_g2_batch(edge, dst)
        for i in range(len(edge)):
            assert e1[i].tobytes() == o.g1_to_bytes(h.hash_to_g1(edge[i], dst))
            assert e2[i].tobytes() == o.g2_to_bytes(h.hash_to_g2(edge[i], dst))
    # (blob, offsets) form of the same call; enough messages for several pipelined chunks
    many = [b"%06d" % i * (i % 5) for i in range(300000)]
    blob = np.frombuffer(b"".join(many), dtype=np.uint8)
    offs = np.concatenate([[0], np.cumsum([len(m) for m in many])]).astype(np.uint64)
    big = engine.hash_to_g1_batch((blob, offs), b"D")
    assert (big == engine.hash_to_g1_batch(many, b"D")).all()
    for i in (0, 1, 4, 147455, 150000, 299999):
        assert big[i].tobytes() == o.g1_to_bytes(h.hash_to_g1(many[i], b"D"))
    with pytest.raises(ValueError):
        engine.hash_to_g1_batch((blob[:10], offs), b"D")
    assert engine.hash_to_g2_batch([], b"x").shape == (0, 128)
    from gopairingbasedcryptography_b200.bn254 import EngineError
    with pytest.raises(EngineError):
        engine.hash_to_g1_batch([b"m"], b"d" * 256)


@pytest.mark.gpu
def test_gpu_bls_sign_verify_with_real_hash(engine):
    """Config 1 end to end (signature/bls01_signature/bls_signature.go:41-89): pk = [sk]g1, sigma = [sk]H(m) with
    H = hash.BytesToG2, verify e(pk, H(m)) e(-g1, sigma) == 1; a tampered message fails."""
    from gopairingbasedcryptography_b200 import schemes

    n = 200
    g1, _ = port.generators()
    sk = 0x0123456789ABCDEF0FEDCBA9876543210F1E2D3C4B5A6978 % o.R
    skb = common.scalar_bytes([sk])
    pk = engine.g1_mul_base_batch(g1, skb)[0]
    msgs = [b"message-%d" % i for i in range(n)]
    hm = schemes.bytes_to_g2_batch(engine, msgs)
    sigma = engine.g2_mul_batch(hm, np.tile(skb, n))
    ok = schemes.bls_verify_batch(engine, pk, schemes.neg_g1(g1)[0], hm, sigma)
    assert ok.all()
    tampered = list(msgs)
    for i in range(0, n, 5):
        tampered[i] = msgs[i] + b"!"
    ok2 = schemes.bls_verify_batch(engine, pk, schemes.neg_g1(g1)[0], schemes.bytes_to_g2_batch(engine, tampered), sigma)
    assert (ok2 == np.array([i % 5 != 0 for i in range(n)])).all()
