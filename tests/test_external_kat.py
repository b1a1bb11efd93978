"""External known-answer test for the pairing: the EIP-197 (alt_bn128 pairing precompile) vector 'jeff1' of go-ethereum's
bn256Pairing test set -- a published input whose PairingCheck is true.  It pins, from outside this repository, the curve
and twist equations, the G2 generator, the group laws and the Miller loop + final exponentiation as a non-degenerate
bilinear map.  (A check against 1 cannot pin the cofactor of the final exponent; that stays the one unpinned item.)"""
import ctypes
import json
import os

import numpy as np
import pytest

from oracle import bn254_ref as o
from oracle import port

from test_emu_device_code import emu  # noqa: F401  (fixture)

HERE = os.path.dirname(os.path.abspath(__file__))


def load():
    with open(os.path.join(HERE, "golden", "eip197_pairing_check.json")) as f:
        v = json.load(f)
    w = [int(v["input"][i * 64:(i + 1) * 64], 16) for i in range(12)]
    P = [(w[0], w[1]), (w[6], w[7])]
    Q = [((w[3], w[2]), (w[5], w[4])), ((w[9], w[8]), (w[11], w[10]))]  # EVM order: imaginary part first
    return v, P, Q


def raw(P, Q):
    return (np.frombuffer(b"".join(o.g1_to_bytes(p) for p in P), dtype=np.uint8).copy(),
            np.frombuffer(b"".join(o.g2_to_bytes(q) for q in Q), dtype=np.uint8).copy())


def test_eip197_vector_definitional_oracle_and_c_port():
    v, P, Q = load()
    assert all(o.g1_on_curve(p) for p in P) and all(o.g2_on_curve(q) for q in Q) and Q[1] == o.G2_GEN
    assert o.pairing_check(P, Q) is v["expected"]
    assert not o.pairing_check([P[0], o.g1_add(P[1], o.G1_GEN)], Q)  # a perturbed input fails
    Pb, Qb = raw(P, Q)
    assert bool(port.pairing_check_batch(Pb, Qb, 1, 2, 1)[0]) is v["expected"]
    assert port.multi_pair_batch(Pb, Qb, 1, 2, 1).tobytes() == o.gt_to_bytes(o.FP12_ONE)
    # e(P1, Q1) alone is not 1, and equals e(-P2, G2): the relation the vector encodes
    e1 = port.pair_batch(Pb[:64], Qb[:128], 1)
    assert e1.tobytes() != o.gt_to_bytes(o.FP12_ONE)
    assert e1.tobytes() == port.pair_batch(np.frombuffer(o.g1_to_bytes(o.g1_neg(P[1])), dtype=np.uint8), Qb[128:], 1).tobytes()


def test_eip197_vector_device_code_on_host(emu):  # noqa: F811
    _, P, Q = load()
    Pb, Qb = raw(P, Q)
    out = np.zeros(1, dtype=np.uint8)
    emu.emu_multi_pair(Pb.ctypes.data_as(ctypes.c_void_p), Qb.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(1), ctypes.c_size_t(2),
                       2, out.ctypes.data_as(ctypes.c_void_p))
    assert out[0] == 1


@pytest.mark.gpu
def test_eip197_vector_gpu(engine):
    from gopairingbasedcryptography_b200 import bn254

    _, P, Q = load()
    Pb, Qb = raw(P, Q)
    for e in (engine, bn254.default_engine()):  # thread kernels and the small-batch lane-group route
        assert e.pairing_check_batch(Pb, Qb, 2)[0]
        assert e.multi_pair_batch(Pb, Qb, 2).tobytes() == o.gt_to_bytes(o.FP12_ONE)
        assert e.pairing_check2_fixed_g1_batch(Pb[:64], Pb[64:], Qb[:128], Qb[128:])[0]
    # 200 copies fill a CTA: the lockstep path
    assert engine.pairing_check_batch(np.tile(Pb, 200), np.tile(Qb, 200), 2).all()


# ---- EIP-196 (alt_bn128 add / scalar-mul precompiles): exact affine outputs, so these pin G1 Add and
# ScalarMultiplication byte for byte -----------------------------------------------------------------------------------
def load196():
    with open(os.path.join(HERE, "golden", "eip197_pairing_check.json")) as f:
        v = json.load(f)["eip196"]
    wm = [int(v["mul"]["input"][i:i + 64], 16) for i in range(0, 192, 64)]
    em = [int(v["mul"]["expected"][i:i + 64], 16) for i in range(0, 128, 64)]
    wa = [int(v["add"]["input"][i:i + 64], 16) for i in range(0, 256, 64)]
    ea = [int(v["add"]["expected"][i:i + 64], 16) for i in range(0, 128, 64)]
    return ((wm[0], wm[1]), wm[2], (em[0], em[1])), ((wa[0], wa[1]), (wa[2], wa[3]), (ea[0], ea[1]))


def g1b(p):
    return np.frombuffer(o.g1_to_bytes(p), dtype=np.uint8).copy()


def test_eip196_vectors_oracles():
    (P, k, R), (A, B, S) = load196()
    assert o.g1_on_curve(P) and o.g1_mul(P, k) == R and o.g1_add(A, B) == S
    kb = np.frombuffer(o.scalar_to_bytes(k), dtype=np.uint8)
    assert port.g1_mul_batch(g1b(P), kb, 1).tobytes() == o.g1_to_bytes(R)
    assert port.g1_add_batch(g1b(A), g1b(B), 1).tobytes() == o.g1_to_bytes(S)


def test_eip196_vectors_device_code_on_host(emu):  # noqa: F811
    (P, k, R), (A, B, S) = load196()
    kb = np.frombuffer(o.scalar_to_bytes(k), dtype=np.uint8).copy()
    out = np.zeros(64, dtype=np.uint8)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    pb, ab, bb = g1b(P), g1b(A), g1b(B)
    emu.emu_g1_mul(vp(pb), ctypes.c_size_t(1), vp(kb), ctypes.c_size_t(1), vp(out))
    assert out.tobytes() == o.g1_to_bytes(R)
    emu.emu_g1_add(vp(ab), vp(bb), ctypes.c_size_t(1), vp(out))
    assert out.tobytes() == o.g1_to_bytes(S)


@pytest.mark.gpu
def test_eip196_vectors_gpu(engine):
    (P, k, R), (A, B, S) = load196()
    kb = np.frombuffer(o.scalar_to_bytes(k), dtype=np.uint8).copy()
    assert engine.g1_mul_batch(g1b(P), kb).tobytes() == o.g1_to_bytes(R)
    assert engine.g1_add_batch(g1b(A), g1b(B)).tobytes() == o.g1_to_bytes(S)
    n = 5000  # >= 4096 scalars on one base: the fixed-base window-table path
    assert (engine.g1_mul_base_batch(g1b(P), np.tile(kb, n)) == np.frombuffer(o.g1_to_bytes(R), dtype=np.uint8)).all()


# ---- gnark-crypto's own HashToG1 known answer (ecc/bn254/hash_vectors_test.go): pins expand_message_xmd, hash_to_field
# (L = 48), the SVDW constants and sign choice, and the G1 addition of the two mapped points, byte for byte ------------
def load_h2c():
    with open(os.path.join(HERE, "golden", "eip197_pairing_check.json")) as f:
        v = json.load(f)["gnark_hash_to_g1"]
    return v["msg"].encode(), v["dst"].encode(), (int(v["x"], 16), int(v["y"], 16))


def test_gnark_hash_to_g1_vector_oracle_and_device_code_on_host(emu):  # noqa: F811
    from oracle import hash_to_curve_ref as h2c

    msg, dst, want = load_h2c()
    assert o.g1_on_curve(want)
    assert h2c.hash_to_g1(msg, dst) == want
    out = (ctypes.c_uint8 * 64)()
    emu.emu_hash_to_g1(msg, ctypes.c_size_t(len(msg)), dst, ctypes.c_size_t(len(dst)), out)
    assert bytes(out) == o.g1_to_bytes(want)


@pytest.mark.gpu
def test_gnark_hash_to_g1_vector_gpu(engine):
    msg, dst, want = load_h2c()
    got = engine.hash_to_g1_batch([msg] * 130 + [b"x"], dst)  # a full CTA (lockstep) and a ragged one
    assert (got[:130] == np.frombuffer(o.g1_to_bytes(want), dtype=np.uint8)).all()
    assert got[130].tobytes() != o.g1_to_bytes(want)
