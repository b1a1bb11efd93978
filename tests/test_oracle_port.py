"""C restatement (oracle/bn254_port.c) pinned bit-for-bit to the definitional Python oracle on seeded
inputs, plus the reference's own scheme-level identities re-run on top of it."""
import numpy as np

from oracle import bn254_ref as o
from oracle import port

import common


def test_port_matches_ref_on_seeded_pairs():
    n = 4
    P, Q, a, b = common.points(n, threads=1)
    out = port.pair_batch(P, Q, n, threads=2)
    for i in range(n):
        Pi, Qi = o.g1_from_bytes(P[64 * i:64 * i + 64].tobytes()), o.g2_from_bytes(Q[128 * i:128 * i + 128].tobytes())
        assert Pi == o.g1_mul(o.G1_GEN, a[i]) and Qi == o.g2_mul(o.G2_GEN, b[i])
        assert out[384 * i:384 * i + 384].tobytes() == o.gt_to_bytes(o.pair([Pi], [Qi]))


def test_port_miller_then_final_exp_composes():
    n = 3
    P, Q, _, _ = common.points(n, seed=77)
    ml = port.miller_loop_batch(P, Q, n, 1)
    assert (port.final_exp_batch(ml, n) == port.pair_batch(P, Q, n)).all()


def test_port_threads_agree():
    n = 16
    P, Q, _, _ = common.points(n, seed=5)
    assert (port.pair_batch(P, Q, n, threads=1) == port.pair_batch(P, Q, n, threads=8)).all()


def test_cyclotomic_square_equals_generic_square_in_subgroup():
    n = 3
    P, Q, _, _ = common.points(n, seed=9)
    gt = port.pair_batch(P, Q, n)
    assert (port.gt_cyclo_sqr_batch(gt, n) == port.gt_sqr_batch(gt, n)).all()


def test_bls_round_trip_on_port():
    """signature/bls01_signature/bls_signature.go:58-89 with H(m) := [h]G2 (synthetic hash)."""
    g1, g2 = port.generators()
    sk, h = 0xA5A5A5A5DEADBEEF % o.R, 0x1337C0DE
    sb = common.scalar_bytes
    pk = port.g1_mul_base_batch(g1, sb([sk]), 1)
    hm = port.g2_mul_base_batch(g2, sb([h]), 1)
    sigma = port.g2_mul_batch(hm, sb([sk]), 1)
    neg_sigma = np.frombuffer(o.g2_to_bytes(o.g2_neg(o.g2_from_bytes(sigma.tobytes()))), dtype=np.uint8)
    P = np.concatenate([pk, g1])
    Q = np.concatenate([hm, neg_sigma])
    assert port.pairing_check_batch(P, Q, 1, 2)[0] == 1
    bad = port.g2_mul_base_batch(g2, sb([h + 1]), 1)
    assert port.pairing_check_batch(P, np.concatenate([bad, neg_sigma]), 1, 2)[0] == 0


def test_edge_scalars_and_bilinearity_checksum():
    ks = common.scalars(12)
    g1, g2 = port.generators()
    out = port.g1_mul_base_batch(g1, common.scalar_bytes(ks), len(ks))
    assert out[:64].tobytes() == bytes(64)  # [0]G = infinity
    assert out[64:128].tobytes() == g1.tobytes()
    assert o.g1_from_bytes(out[192:256].tobytes()) == o.g1_neg(o.G1_GEN)  # [r-1]G = -G
    # prod e(a_i G1, b_i G2) == e(G1,G2)^(sum a_i b_i)
    n = 6
    P, Q, a, b = common.points(n, seed=3)
    lhs = port.multi_pair_batch(P, Q, 1, n)
    e = port.pair_batch(g1, g2, 1)
    s = sum(x * y for x, y in zip(a, b)) % o.R
    assert (lhs == port.gt_exp_batch(e, common.scalar_bytes([s]), 1)).all()
