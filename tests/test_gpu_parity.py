"""Parity proper: the CUDA path, called through the C ABI, against the oracle on the same seeded
inputs -- bit-exact (integer work).  Run with `-m gpu` on a B200."""
import json
import os

import numpy as np
import pytest

from oracle import bn254_ref as o
from oracle import port

import common

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def hx(s):
    return np.frombuffer(bytes.fromhex(s), dtype=np.uint8)


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(HERE, "golden", "bn254_vectors.json")) as f:
        return json.load(f)


def test_fp_mul(engine):
    rng = o.SplitMix64(1)
    n = 1 << 14
    a = np.frombuffer(b"".join(rng.fp().to_bytes(32, "little") for _ in range(n)), dtype=np.uint8)
    b = np.frombuffer(b"".join(rng.fp().to_bytes(32, "little") for _ in range(n)), dtype=np.uint8)
    assert (engine.fp_mul_batch(a, b).reshape(-1) == port.fp_mul_batch(a, b, n, 8)).all()


def test_golden_vectors(engine, golden):
    for v in golden["pair"]:
        assert engine.pair_batch(hx(v["P"]), hx(v["Q"])).tobytes().hex() == v["gt"]
    for v in golden["multi_pair"]:
        assert engine.multi_pair_batch(hx(v["P"]), hx(v["Q"]), v["k"]).tobytes().hex() == v["gt"]
    v = golden["final_exp"]
    assert engine.final_exp_batch(hx(v["in"])).tobytes().hex() == v["out"]
    for v in golden["g1_mul"]:
        assert engine.g1_mul_batch(hx(v["base"]), hx(v["k"])).tobytes().hex() == v["out"]
    for v in golden["g2_mul"]:
        assert engine.g2_mul_batch(hx(v["base"]), hx(v["k"])).tobytes().hex() == v["out"]
    for v in golden["g1_add"]:
        assert engine.g1_add_batch(hx(v["a"]), hx(v["b"])).tobytes().hex() == v["out"]
    for v in golden["g2_add"]:
        assert engine.g2_add_batch(hx(v["a"]), hx(v["b"])).tobytes().hex() == v["out"]
    for v in golden["gt_exp"]:
        assert engine.gt_exp_batch(hx(v["x"]), hx(v["k"])).tobytes().hex() == v["out"]
    v = golden["gt_mul"]
    assert engine.gt_mul_batch(hx(v["a"]), hx(v["b"])).tobytes().hex() == v["mul"]
    assert engine.gt_div_batch(hx(v["a"]), hx(v["b"])).tobytes().hex() == v["div"]


def test_pair_batch_vs_oracle(engine):
    n = 1024 + 7  # ragged: not a multiple of the block size
    P, Q, _, _ = common.points(n, threads=8)
    P, Q = common.with_infinities(P, Q)
    out = engine.pair_batch(P, Q)
    assert (out.reshape(-1) == port.pair_batch(P, Q, n, 8)).all()
    one = o.gt_to_bytes(o.FP12_ONE)
    assert out[0].tobytes() == one and out[1].tobytes() == one and out[2].tobytes() == one


def test_pair_batch_vs_oracle_20k(engine):
    """A larger bit-exact sweep of the throughput kernels (lockstep CTAs, many distinct operands) against the C port."""
    n = 20000 + 37
    P, Q, _, _ = common.points(n, seed=4321, threads=16)
    assert (engine.pair_batch(P, Q).reshape(-1) == port.pair_batch(P, Q, n, 16)).all()
    ml = engine.miller_loop_batch(P, Q, 1)
    assert (engine.final_exp_batch(ml).reshape(-1) == port.pair_batch(P, Q, n, 16)).all()


def test_single_element_and_errors(engine):
    P, Q, _, _ = common.points(1, seed=99)
    assert (engine.pair_batch(P, Q).reshape(-1) == port.pair_batch(P, Q, 1)).all()
    with pytest.raises(ValueError, match="invalid inputs sizes"):
        engine.pair_batch(P, np.concatenate([Q, Q]))
    with pytest.raises(ValueError, match="invalid inputs sizes"):
        engine.multi_pair_batch(P, Q, 0)
    assert engine.pair_batch(b"", b"").shape == (0, 384)


@pytest.mark.parametrize("k", [2, 3, 5, 9])
def test_multi_pair_and_check(engine, k):
    n = 48
    P, Q, _, _ = common.points(n * k, seed=1000 + k, threads=8)
    P, Q = common.with_infinities(P, Q)
    assert (engine.multi_pair_batch(P, Q, k).reshape(-1) == port.multi_pair_batch(P, Q, n, k, 8)).all()
    ml = engine.miller_loop_batch(P, Q, k)
    assert (engine.final_exp_batch(ml).reshape(-1) == port.multi_pair_batch(P, Q, n, k, 8)).all()
    assert (engine.pairing_check_batch(P, Q, k) == port.pairing_check_batch(P, Q, n, k, 8).astype(bool)).all()


def test_split_multi_pair_large_k(engine):
    """k > 16 takes the split path (groups of 8 pairs per thread, then combine + one final exp)."""
    n, k = 6, 37
    P, Q, _, _ = common.points(n * k, seed=2024, threads=8)
    P, Q = common.with_infinities(P, Q)
    ref = port.multi_pair_batch(P, Q, n, k, 8)
    assert (engine.multi_pair_batch(P, Q, k).reshape(-1) == ref).all()
    assert (engine.final_exp_batch(engine.miller_loop_batch(P, Q, k)).reshape(-1) == ref).all()
    # e(P,Q)^-1 appended k times makes the product 1 for item 0 only
    ok = engine.pairing_check_batch(P, Q, k)
    assert (ok == port.pairing_check_batch(P, Q, n, k, 8).astype(bool)).all()


def test_full_cta_lockstep_paths(engine):
    """Full CTAs of 128 elements take the barrier-synchronised (lockstep) path of the pairing kernels; CTAs that are
    ragged or hold a point at infinity fall back to the free-running one.  Both must agree with the oracle, including
    elements whose final exponentiation leaves the easy part at 1 (gnark's early return) next to ones that do not."""
    for k, n in ((1, 256 + 3), (2, 256), (3, 128 + 1), (5, 128 + 2), (9, 128), (37, 128 + 1)):
        P, Q, _, _ = common.points(n * k, seed=7000 + k, threads=8)
        P, Q = P.copy(), Q.copy()
        if k >= 2:
            # element 5: e(A,B) e(-A,B) -> the Miller product dies in the easy part; element n-1: infinity member
            A, B = P[:64].copy(), Q[:128].copy()
            negA = A.copy()
            y = int.from_bytes(A[32:64].tobytes(), "little")
            negA[32:64] = np.frombuffer(((o.P - y) % o.P).to_bytes(32, "little"), dtype=np.uint8)
            base = 5 * k
            P[64 * base:64 * (base + 1)] = A
            Q[128 * base:128 * (base + 1)] = B
            P[64 * (base + 1):64 * (base + 2)] = negA
            Q[128 * (base + 1):128 * (base + 2)] = B
            for j in range(2, k):
                P[64 * (base + j):64 * (base + j + 1)] = 0
            P[64 * ((n - 1) * k):64 * ((n - 1) * k + 1)] = 0
        ref = port.multi_pair_batch(P, Q, n, k, 8) if k > 1 else port.pair_batch(P, Q, n, 8)
        got = engine.multi_pair_batch(P, Q, k) if k > 1 else engine.pair_batch(P, Q)
        assert (got.reshape(-1) == ref).all(), k
        if k > 1:
            assert got[5].tobytes() == o.gt_to_bytes(o.FP12_ONE)
            assert (engine.final_exp_batch(engine.miller_loop_batch(P, Q, k)).reshape(-1) == ref).all(), k
            assert (engine.pairing_check_batch(P, Q, k) == port.pairing_check_batch(P, Q, n, k, 8).astype(bool)).all(), k
    # final exponentiation alone on a full CTA holding 1, 0 and random Fp12 values
    rng = o.SplitMix64(31337)
    vals = [o.FP12_ONE] + [tuple(tuple((rng.fp(), rng.fp()) for _ in range(3)) for _ in range(2)) for _ in range(127)]
    buf = np.frombuffer(b"".join(o.gt_to_bytes(v) for v in vals), dtype=np.uint8).copy()
    buf[384:768] = 0
    assert (engine.final_exp_batch(buf).reshape(-1) == port.final_exp_batch(buf, 128, 8)).all()


def test_fixed_base_tables(engine):
    """>= 4096 scalars on one base switch to the cached 32x255 window table; the cache follows the base."""
    n = 4096 + 5
    ks = common.scalars(n) + [(1 << 256) - 1, o.R, o.R + 1]
    n = len(ks)
    sb = common.scalar_bytes(ks)
    g1, g2 = port.generators()
    assert (engine.g1_mul_base_batch(g1, sb).reshape(-1) == port.g1_mul_base_batch(g1, sb, n, 8)).all()
    assert (engine.g2_mul_base_batch(g2, sb).reshape(-1) == port.g2_mul_base_batch(g2, sb, n, 8)).all()
    P, Q, _, _ = common.points(2, seed=606)
    assert (engine.g1_mul_base_batch(P[64:128], sb).reshape(-1) == port.g1_mul_base_batch(P[64:128], sb, n, 8)).all()
    assert (engine.g1_mul_base_batch(g1, sb).reshape(-1) == port.g1_mul_base_batch(g1, sb, n, 8)).all()
    assert (engine.g2_mul_base_batch(Q[:128], sb[:32 * 4100]).reshape(-1) == port.g2_mul_base_batch(Q[:128], sb[:32 * 4100], 4100, 8)).all()


def test_bls_verify_batch(engine):
    """Config 1 shape (signature/bls01_signature/bls_signature.go:58-89), H(m) := [h_i]G2 synthetic hash:
    sign on the GPU, verify on the GPU, flip some messages -> those verify false."""
    n = 256
    g1, g2 = port.generators()
    sk = 0x1F2E3D4C5B6A79880123456789ABCDEF % o.R
    hs = common.scalars(n, seed=4242, edges=False)
    skb = common.scalar_bytes([sk])
    pk = engine.g1_mul_base_batch(g1, skb)
    hm = engine.g2_mul_base_batch(g2, common.scalar_bytes(hs))
    sigma = engine.g2_mul_batch(hm, np.tile(skb, n))
    assert (sigma.reshape(-1) == port.g2_mul_batch(hm.reshape(-1), np.tile(skb, n), n, 8)).all()
    neg = sigma.copy().reshape(n, 4, 32)
    for i in range(n):
        for c in (2, 3):
            v = int.from_bytes(neg[i, c].tobytes(), "little")
            neg[i, c] = np.frombuffer(((o.P - v) % o.P).to_bytes(32, "little"), dtype=np.uint8)
    hm_bad = hm.copy()
    bad = list(range(0, n, 4))
    hm_bad[bad] = engine.g2_mul_base_batch(g2, common.scalar_bytes([hs[i] + 1 for i in bad]))
    P = np.concatenate([np.tile(pk.reshape(1, 64), (n, 1)), np.tile(g1.reshape(1, 64), (n, 1))], axis=1)  # (n, 2*64)
    Qg = np.concatenate([hm_bad, neg.reshape(n, 128)], axis=1)
    ok = engine.pairing_check_batch(P, Qg, 2)
    expect = np.array([i not in bad for i in range(n)])
    assert (ok == expect).all()
    assert (ok == port.pairing_check_batch(P, Qg, n, 2, 8).astype(bool)).all()


def test_groups_vs_oracle(engine):
    n = 300
    ks = common.scalars(n) + [(1 << 256) - 1]
    n += 1
    sb = common.scalar_bytes(ks)
    P, Q, _, _ = common.points(n, seed=31, threads=8)
    P[64 * 9:64 * 10] = 0  # infinity base
    assert (engine.g1_mul_batch(P, sb).reshape(-1) == port.g1_mul_batch(P, sb, n, 8)).all()
    assert (engine.g2_mul_batch(Q, sb).reshape(-1) == port.g2_mul_batch(Q, sb, n, 8)).all()
    g1, g2 = port.generators()
    assert (engine.g1_mul_base_batch(g1, sb).reshape(-1) == port.g1_mul_base_batch(g1, sb, n, 8)).all()
    assert (engine.g2_mul_base_batch(g2, sb).reshape(-1) == port.g2_mul_base_batch(g2, sb, n, 8)).all()
    Pr, Qr = np.roll(P, 64), np.roll(Q, 128)
    assert (engine.g1_add_batch(P, Pr).reshape(-1) == port.g1_add_batch(P, Pr, n, 8)).all()
    assert (engine.g1_add_batch(P, P).reshape(-1) == port.g1_add_batch(P, P, n, 8)).all()
    assert (engine.g2_add_batch(Q, Qr).reshape(-1) == port.g2_add_batch(Q, Qr, n, 8)).all()


def test_gt_vs_oracle(engine):
    n = 64
    P, Q, _, _ = common.points(n, seed=55, threads=8)
    gt = engine.pair_batch(P, Q)
    sb = common.scalar_bytes(common.scalars(n))
    assert (engine.gt_exp_batch(gt, sb).reshape(-1) == port.gt_exp_batch(gt.reshape(-1), sb, n, 8)).all()
    assert (engine.gt_exp_base_batch(gt[0], sb).reshape(-1) == port.gt_exp_base_batch(gt[0], sb, n, 8)).all()
    gt2 = np.roll(gt, 1, axis=0)
    assert (engine.gt_mul_batch(gt, gt2).reshape(-1) == port.gt_mul_batch(gt.reshape(-1), gt2.reshape(-1), n, 8)).all()
    assert (engine.gt_div_batch(gt, gt2).reshape(-1) == port.gt_div_batch(gt.reshape(-1), gt2.reshape(-1), n, 8)).all()


@pytest.mark.parametrize("log2n", [14, 20])
def test_bilinearity_checksum_large(engine, log2n):
    """Size-independent property at a size the oracle cannot finish (2^20 = BASELINE configs[1]):
    prod_i e(a_i G1, b_i G2) == e(G1,G2)^(sum a_i b_i), with the points generated on the GPU."""
    n = 1 << log2n
    a = common.scalars(n, seed=7, edges=False)
    b = common.scalars(n, seed=8, edges=False)
    g1, g2 = port.generators()
    P = engine.g1_mul_base_batch(g1, common.scalar_bytes(a))
    Q = engine.g2_mul_base_batch(g2, common.scalar_bytes(b))
    gt = engine.pair_batch(P, Q)
    # sampled bit-exact check
    idx = [0, 1, n // 2, n - 1]
    ref = port.pair_batch(P[idx].reshape(-1), Q[idx].reshape(-1), len(idx), 4).reshape(len(idx), 384)
    assert (gt[idx] == ref).all()
    # product tree on the GPU
    cur = gt
    while cur.shape[0] > 1:
        h = cur.shape[0] // 2
        cur = engine.gt_mul_batch(cur[:h], cur[h:2 * h])
    s = sum(x * y for x, y in zip(a, b)) % o.R
    e = port.pair_batch(g1, g2, 1)
    assert cur[0].tobytes() == port.gt_exp_batch(e, common.scalar_bytes([s]), 1).tobytes()


def test_gnark_named_api(engine):
    """Reads like the reference's tests (e.g. signature/zss04_signature/zss04_signature_test.go:26-38)."""
    from gopairingbasedcryptography_b200 import bn254

    _, _, g1, g2 = bn254.Generators()
    e1 = bn254.Pair([g1], [g2])
    e2 = bn254.Pair([g1], [g2])
    assert e1 == e2 and e1.Equal(e2)  # deterministic
    with pytest.raises(ValueError, match="invalid inputs sizes"):
        bn254.Pair([], [])
    x = 0x1234567
    pk = bn254.G1Affine().ScalarMultiplicationBase(x)
    hm = bn254.G2Affine().ScalarMultiplicationBase(987654321)
    sig = bn254.G2Affine().ScalarMultiplication(hm, x)
    assert bn254.PairingCheck([pk, g1], [hm, bn254.G2Affine().Neg(sig)])
    assert not bn254.PairingCheck([pk, g1], [hm, sig])
    assert bn254.FinalExponentiation(bn254.MillerLoop([pk], [hm])) == bn254.Pair([pk], [hm])
    # GT.Exp / Mul / Div / Inverse and negative exponents
    a = bn254.GT().Exp(e1, 5)
    b = bn254.GT().Exp(e1, -5)
    assert bn254.GT().Mul(a, b) == bn254.GT().SetOne()
    assert bn254.GT().Div(a, a) == bn254.GT().SetOne()
    assert bn254.GT().Exp(e1, 0) == bn254.GT().SetOne()
    assert bn254.Pair([pk], [hm]) == bn254.GT().Exp(bn254.Pair([g1], [hm]), x)
    # negative / oversized group scalars follow big.Int semantics
    assert bn254.G1Affine().ScalarMultiplication(g1, -1) == bn254.G1Affine().Neg(g1)
    assert bn254.G1Affine().ScalarMultiplication(g1, bn254.R_MOD + 2) == bn254.G1Affine().Add(g1, g1)
    assert bn254.G1Affine().Sub(g1, g1).IsInfinity()


def test_device_pointer_entry_points(engine):
    import torch

    n = 512
    P, Q, _, _ = common.points(n, seed=77, threads=8)
    dP = torch.from_numpy(P.copy()).cuda()
    dQ = torch.from_numpy(Q.copy()).cuda()
    dO = torch.empty(n * 384, dtype=torch.uint8, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    engine.pair_batch_dev(dP.data_ptr(), dQ.data_ptr(), n, dO.data_ptr(), s)
    torch.cuda.synchronize()
    assert (dO.cpu().numpy() == port.pair_batch(P, Q, n, 8)).all()
    dM = torch.empty(n * 384, dtype=torch.uint8, device="cuda")
    engine.miller_loop_batch_dev(dP.data_ptr(), dQ.data_ptr(), n, 1, dM.data_ptr(), s)
    engine.final_exp_batch_dev(dM.data_ptr(), n, dM.data_ptr(), s)
    torch.cuda.synchronize()
    assert torch.equal(dM, dO)


def test_bsw07_fused_decrypt_matches_unfused_reference_formula(engine):
    """Fused (2m+1)-pair product + one final exp vs the reference's unfused flow
    (access/tree/access_tree_node.go:96-164: per leaf Pair/Pair/Div, GT.Exp by the Lagrange coefficient, Mul;
    cpabe/bsw07/bsw07_cpabe.go:184-190: Pair, Div, Div), computed with the oracle."""
    from gopairingbasedcryptography_b200 import schemes

    n, m = 3, 5
    cyP, djQ, _, _ = common.points(n * m, seed=71, threads=8)
    cypP, djpQ, _, _ = common.points(n * m, seed=72, threads=8)
    cP, dQ, _, _ = common.points(n, seed=73)
    cy = cyP.reshape(n, m, 64)
    cyp = cypP.reshape(n, m, 64)
    dj = djQ.reshape(n * m, 128)[:m]
    djp = djpQ.reshape(n * m, 128)[:m]
    d = dQ[:128]
    c = cP.reshape(n, 64)
    deltas = common.scalar_bytes(common.scalars(m, seed=74, edges=False)).reshape(m, 32)
    ctil = engine.pair_batch(c, np.tile(d, n))  # any GT elements
    got = schemes.bsw07_decrypt_batch(engine, cy, cyp, dj, djp, c, d, ctil, deltas)
    for i in range(n):
        A = None
        for j in range(m):
            e1 = port.pair_batch(cy[i, j], dj[j], 1)
            e2 = port.pair_batch(cyp[i, j], djp[j], 1)
            fz = port.gt_exp_batch(port.gt_div_batch(e1, e2, 1), deltas[j], 1)
            A = fz if A is None else port.gt_mul_batch(A, fz, 1)
        ecd = port.pair_batch(c[i], d, 1)
        M = port.gt_div_batch(ctil[i], port.gt_div_batch(ecd, A, 1), 1)
        assert (got[i] == M).all()
    # Lagrange coefficients folded into the key's line tables: no per-ciphertext scalar multiplication
    pol = schemes.bsw07_policy_lines(engine, dj, djp, d, deltas)
    assert (schemes.bsw07_decrypt_batch(engine, cy, cyp, dj, djp, c, d, ctil, deltas, lines=pol, folded=True) == got).all()


def test_bls_verify_driver(engine):
    from gopairingbasedcryptography_b200 import schemes

    n = 64
    g1, g2 = port.generators()
    sk = 12345678901234567890
    skb = common.scalar_bytes([sk])
    pk = engine.g1_mul_base_batch(g1, skb)[0]
    hm = engine.g2_mul_base_batch(g2, common.scalar_bytes(common.scalars(n, seed=9, edges=False)))
    sigma = engine.g2_mul_batch(hm, np.tile(skb, n))
    neg = sigma.copy().reshape(n, 4, 32)
    for i in range(n):
        for cidx in (2, 3):
            v = int.from_bytes(neg[i, cidx].tobytes(), "little")
            neg[i, cidx] = np.frombuffer(((o.P - v) % o.P).to_bytes(32, "little"), dtype=np.uint8)
    ok = schemes.bls_verify_batch(engine, pk, g1, hm, neg.reshape(n, 128))
    assert ok.all()
    hm2 = hm.copy()
    hm2[5] = hm[6]
    assert not schemes.bls_verify_batch(engine, pk, g1, hm2, neg.reshape(n, 128))[5]


def test_gt_cyclotomic_exp(engine):
    n = 40
    P, Q, _, _ = common.points(n, seed=56, threads=8)
    gt = engine.pair_batch(P, Q)
    ks = common.scalars(n - 4) + [(1 << 256) - 1, 1 << 255, 3, 4]
    sb = common.scalar_bytes(ks)
    ref = port.gt_exp_batch(gt.reshape(-1), sb, n, 8)
    assert (engine.gt_cyclo_exp_batch(gt, sb).reshape(-1) == ref).all()
    assert (engine.gt_exp_batch(gt, sb).reshape(-1) == ref).all()
    assert (engine.gt_cyclo_exp_base_batch(gt[3], sb).reshape(-1) == port.gt_exp_base_batch(gt[3], sb, n, 8)).all()


def test_subset_and_segment_sums(engine):
    m, n = 256, 33
    U1, U2, _, _ = common.points(m + 1, seed=91, threads=8)
    rng = np.random.default_rng(5)
    sel = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    sel[0] = 0
    sel[1] = 255
    got2 = engine.g2_subset_sum_batch(U2, sel)
    got1 = engine.g1_subset_sum_batch(U1, sel)
    for i in (0, 1, 2, n - 1):
        a1, a2 = U1[:64].copy(), U2[:128].copy()
        for j in range(m):
            if (sel[i, j >> 3] >> (7 - (j & 7))) & 1:
                a1 = port.g1_add_batch(a1, U1[64 * (j + 1):64 * (j + 2)], 1)
                a2 = port.g2_add_batch(a2, U2[128 * (j + 1):128 * (j + 2)], 1)
        assert (got1[i] == a1).all() and (got2[i] == a2).all()
    # segment sums incl. infinity members, a cancelling pair and non-power-of-two lengths
    for length in (1, 5, 37, 100):
        groups = 7
        pts, qts, _, _ = common.points(groups * length, seed=300 + length, threads=8)
        pts[:64] = 0
        if length >= 5:
            neg = o.g1_to_bytes(o.g1_neg(o.g1_from_bytes(pts[64 * 2:64 * 3].tobytes())))
            pts[64 * 3:64 * 4] = np.frombuffer(neg, dtype=np.uint8)
        s1 = engine.g1_sum_batch(pts, length)
        s2 = engine.g2_sum_batch(qts, length)
        for g in (0, groups - 1):
            a1, a2 = np.zeros(64, np.uint8), np.zeros(128, np.uint8)
            for j in range(length):
                k = g * length + j
                a1 = port.g1_add_batch(a1, pts[64 * k:64 * k + 64], 1)
                a2 = port.g2_add_batch(a2, qts[128 * k:128 * k + 128], 1)
            assert (s1[g] == a1).all() and (s2[g] == a2).all()


def test_waters05_encrypt_decrypt_round_trip(engine):
    """Config 4 shape (ibe/waters05_ibe/waters05_ibe.go:206-279) on the batch entry points: fixed-base G1,
    GT exponentiation of the constant e(g1^alpha, g2), Waters hash as one subset sum, G2 variable-base mult;
    decryption M = c1 * e(d2, c3) / e(c2, d1) recovers the message for every identity."""
    import hashlib
    from gopairingbasedcryptography_b200 import schemes

    n, m = 48, 256
    g1, g2 = port.generators()
    sb = common.scalar_bytes
    alpha = 0x1234567890ABCDEF1234567890ABCDEF % o.R
    us = common.scalars(m + 1, seed=505, edges=False)
    U = engine.g2_mul_base_batch(g2, sb(us))  # U', U_1..U_m
    g1a = engine.g1_mul_base_batch(g1, sb([alpha]))[0]
    e_const = engine.pair_batch(g1a, g2)[0]
    ids = np.stack([np.frombuffer(hashlib.sha256(b"id-%d" % i).digest(), dtype=np.uint8) for i in range(n)])
    H = engine.g2_subset_sum_batch(U, ids)
    ts = sb(common.scalars(n, seed=506, edges=False)).reshape(n, 32)
    rs = sb(common.scalars(n, seed=507, edges=False)).reshape(n, 32)
    msgs = engine.gt_cyclo_exp_base_batch(e_const, sb(common.scalars(n, seed=508, edges=False)))
    # encrypt
    c1 = engine.gt_mul_batch(engine.gt_cyclo_exp_base_batch(e_const, ts), msgs)
    c2 = engine.g1_mul_base_batch(g1, ts)
    c3 = engine.g2_mul_batch(H, ts)
    # key generation: d1 = [alpha]g2 + [r]H(id), d2 = [r]g1
    d1 = engine.g2_add_batch(np.tile(engine.g2_mul_base_batch(g2, sb([alpha])), (n, 1)), engine.g2_mul_batch(H, rs))
    d2 = engine.g1_mul_base_batch(g1, rs)
    # decrypt: one 2-pair product e(d2, c3) * e(-c2, d1) per ciphertext
    P = np.concatenate([d2.reshape(n, 1, 64), schemes.neg_g1(c2).reshape(n, 1, 64)], axis=1)
    Q = np.concatenate([c3.reshape(n, 1, 128), d1.reshape(n, 1, 128)], axis=1)
    rec = engine.gt_mul_batch(c1, engine.multi_pair_batch(P, Q, 2))
    assert (rec == msgs).all()
    # the Waters hash agrees with the reference's Add loop (oracle) for one identity
    acc = U[0].copy()
    for j in range(m):
        if (ids[7, j >> 3] >> (7 - (j & 7))) & 1:
            acc = port.g2_add_batch(acc, U[j + 1], 1)
    assert (H[7] == acc).all()


def test_afp25_shaped_msm(engine):
    """Config 5 shape (bibe/afp25_bibe/afp25_bibe_utils.go:45-55): result = sum_j [coef_j] tauPowers_j for several
    coefficient vectors over the SAME points: n*len variable-base mults + segment sums."""
    length, n = 64, 5
    pts, _, _, _ = common.points(length, seed=808, threads=8)
    coef = common.scalar_bytes(common.scalars(n * length, seed=809, edges=False))
    terms = engine.g1_mul_batch(np.tile(pts, n), coef)
    got = engine.g1_sum_batch(terms, length)
    for i in (0, n - 1):
        acc = np.zeros(64, np.uint8)
        t = port.g1_mul_batch(pts, coef[32 * i * length:32 * (i + 1) * length], length, 8)
        for j in range(length):
            acc = port.g1_add_batch(acc, t[64 * j:64 * j + 64], 1)
        assert (got[i] == acc).all()


def test_precomputed_g2_line_tables(engine):
    """bn254_g2_lines_create + bn254_multi_pair_lines_batch == bn254_multi_pair_batch, bit for bit; the BSW07 driver
    gives the same plaintexts with and without the key's line tables."""
    from gopairingbasedcryptography_b200 import schemes

    for n, m in ((5, 3), (130, 21)):
        _, Q, _, _ = common.points(m, seed=411 + m, threads=8)
        P, _, _, _ = common.points(n * m, seed=412 + m, threads=8)
        Q[128:256] = 0
        P[64 * 4:64 * 5] = 0
        lines = engine.g2_lines_create(Q)
        got = engine.multi_pair_lines_batch(P, lines)
        assert (got.reshape(-1) == port.multi_pair_batch(P, np.tile(Q, n), n, m, 8)).all()
        assert (got == engine.multi_pair_batch(P, np.tile(Q, n), m)).all()
        lines.close()
    n, m = 4, 6
    cyP, djQ, _, _ = common.points(n * m, seed=81, threads=8)
    cypP, djpQ, _, _ = common.points(n * m, seed=82, threads=8)
    cP, dQ, _, _ = common.points(n, seed=83)
    cy, cyp = cyP.reshape(n, m, 64), cypP.reshape(n, m, 64)
    dj, djp, d, c = djQ.reshape(-1, 128)[:m], djpQ.reshape(-1, 128)[:m], dQ[:128], cP.reshape(n, 64)
    deltas = common.scalar_bytes(common.scalars(m, seed=84, edges=False)).reshape(m, 32)
    ctil = engine.pair_batch(c, np.tile(d, n))
    a = schemes.bsw07_decrypt_batch(engine, cy, cyp, dj, djp, c, d, ctil, deltas)
    key = schemes.bsw07_key_lines(engine, dj, djp, d)
    b = schemes.bsw07_decrypt_batch(engine, cy, cyp, dj, djp, c, d, ctil, deltas, lines=key)
    assert (a == b).all()


def test_line_tables_through_the_shared_memory_ring(engine):
    """Full, skip-free CTAs of the line-table Miller kernel fetch every couple of lines with one bulk copy into a
    shared-memory ring (k_pairing.cu); chunk sizes 8 + 1 (single lines only), 8 + 8 + 5 (odd tail) and 2 (one couple per
    position), full CTAs next to a ragged one -- all bit-identical to the generic multi-pairing and to the oracle."""
    for n, m in ((259, 9), (384, 2), (128, 21)):
        _, Q, _, _ = common.points(m, seed=511 + m, threads=8)
        P, _, _, _ = common.points(n * m, seed=512 + m, threads=8)
        lines = engine.g2_lines_create(Q)
        got = engine.multi_pair_lines_batch(P, lines)
        assert (got == engine.multi_pair_batch(P, np.tile(Q, n), m)).all()
        k = 16  # oracle on the first and last items
        for lo in (0, n - k):
            want = port.multi_pair_batch(P[64 * m * lo:64 * m * (lo + k)], np.tile(Q, k), k, m, 8)
            assert (got[lo:lo + k].reshape(-1) == want).all()
        lines.close()


def test_gt_fixed_base_table(engine):
    """>= 4096 exponents on one base use the cached 32x255 GT window table (built with the generic ladder, so it is
    valid for ANY Fp12 base, not only pairing outputs)."""
    n = 4096 + 3
    ks = common.scalars(n - 3) + [0, (1 << 256) - 1, 1 << 255]
    sb = common.scalar_bytes(ks)
    P, Q, _, _ = common.points(2, seed=515)
    gt = engine.pair_batch(P, Q)
    idx = [0, 1, 2, 3, 4, 5, 6, n - 3, n - 2, n - 1, 777]
    sel = np.concatenate([sb[32 * i:32 * i + 32] for i in idx])
    out = engine.gt_cyclo_exp_base_batch(gt[0], sb)
    assert (out[idx].reshape(-1) == port.gt_exp_base_batch(gt[0], sel, len(idx), 8)).all()
    rng = o.SplitMix64(4)
    x = np.frombuffer(b"".join(o.fp_to_mont_bytes(rng.fp()) for _ in range(12)), dtype=np.uint8).copy()
    out = engine.gt_exp_base_batch(x, sb)          # generic element: the cache must switch bases
    assert (out[idx].reshape(-1) == port.gt_exp_base_batch(x, sel, len(idx), 8)).all()
    out = engine.gt_exp_base_batch(gt[1], sb)
    assert (out[idx].reshape(-1) == port.gt_exp_base_batch(gt[1], sel, len(idx), 8)).all()


def test_lane_group_vm_implementation():
    """The alternative pairing implementation (K=3 lane groups, tower VM; BN254_IMPL=vm, read at context creation)
    through the same C ABI: pair / miller / final-exp parity incl. ragged sizes and infinity operands."""
    from gopairingbasedcryptography_b200 import bn254

    old = os.environ.get("BN254_IMPL")
    os.environ["BN254_IMPL"] = "vm"
    try:
        eng = bn254.Engine(0)
    finally:
        if old is None:
            del os.environ["BN254_IMPL"]
        else:
            os.environ["BN254_IMPL"] = old
    n = 10 * 13 + 7  # not a multiple of the 10 pairings a warp owns
    P, Q, _, _ = common.points(n, seed=909, threads=8)
    P, Q = common.with_infinities(P, Q)
    ref = port.pair_batch(P, Q, n, 8)
    assert (eng.pair_batch(P, Q).reshape(-1) == ref).all()
    ml = eng.miller_loop_batch(P, Q, 1)
    assert (eng.final_exp_batch(ml).reshape(-1) == ref).all()
    rng = o.SplitMix64(12)
    x = np.frombuffer(b"".join(o.fp_to_mont_bytes(rng.fp()) for _ in range(24)), dtype=np.uint8).copy()
    assert (eng.final_exp_batch(x).reshape(-1) == port.final_exp_batch(x, 2)).all()
    eng.close()


def test_error_contract(engine):
    """gnark's only error ("invalid inputs sizes") and the ABI's argument checks; nothing else validates input."""
    import ctypes
    from gopairingbasedcryptography_b200 import bn254, _native

    P, Q, _, _ = common.points(2, seed=5)
    with pytest.raises(ValueError, match="invalid inputs sizes"):
        engine.multi_pair_batch(P, Q[:128], 2)
    with pytest.raises(ValueError, match="invalid inputs sizes"):
        engine.pairing_check_batch(P, Q, 0)
    with pytest.raises(ValueError, match="invalid inputs sizes"):
        bn254.PairingCheck([bn254.G1Affine()], [])
    lib = _native.lib()
    lib.bn254_pair_batch.restype = ctypes.c_int
    out = np.zeros(384, np.uint8)
    rc = lib.bn254_pair_batch(engine.handle, None, None, ctypes.c_size_t(1), out.ctypes.data_as(ctypes.c_void_p))
    assert rc == -4 and b"null" in lib.bn254_last_error(engine.handle)
    lib.bn254_multi_pair_batch.restype = ctypes.c_int
    rc = lib.bn254_multi_pair_batch(engine.handle, P.ctypes.data_as(ctypes.c_void_p), Q.ctypes.data_as(ctypes.c_void_p),
                                    ctypes.c_size_t(1), ctypes.c_size_t(0), out.ctypes.data_as(ctypes.c_void_p))
    assert rc == -1
    # points that are not on the curve are processed without complaint, like gnark (no validation): just no crash
    junk = np.arange(64, dtype=np.uint8)
    engine.pair_batch(junk, Q[:128])
    # a second context on the same GPU works concurrently with the first
    e2 = bn254.Engine(0)
    assert (e2.pair_batch(P, Q) == engine.pair_batch(P, Q)).all()
    e2.close()


def test_sw05_bb04_gwww25_fused_drivers(engine):
    """§8f-3 callers: SW05 FIBE (fibe/sw05_fibe_common.go:305-323), BB04-IBE (ibe/bb04_ibe/bb04_ibe.go:213-236) and the
    GWWW25 G2-side MSM (bibe/gwww25_bibe/gwww25_bibe_utils.go:40-50): fused batch drivers vs the reference's unfused
    formulas evaluated with the oracle."""
    from gopairingbasedcryptography_b200 import schemes

    # SW05: M = e' / prod_i e(D_i, E_i)^Delta_i
    n, m = 3, 4
    D, E, _, _ = common.points(n * m, seed=601, threads=8)
    di, ei = D.reshape(n, m, 64), E.reshape(n, m, 128)
    deltas = common.scalar_bytes(common.scalars(n * m, seed=602, edges=False)).reshape(n, m, 32)
    P0, Q0, _, _ = common.points(n, seed=603)
    eprime = engine.pair_batch(P0, Q0)
    got = schemes.sw05_fibe_decrypt_batch(engine, di, ei, eprime, deltas)
    for i in range(n):
        den = None
        for j in range(m):
            t = port.gt_exp_batch(port.pair_batch(di[i, j], ei[i, j], 1), deltas[i, j], 1)
            den = t if den is None else port.gt_mul_batch(den, t, 1)
        assert (got[i] == port.gt_div_batch(eprime[i], den, 1)).all()
    # BB04-IBE: M = a * prod_j e(d_j, c_j) / e(b, d0), one key for the batch, k = 9 here (256 in the reference)
    n, k = 3, 9
    dj, _, _, _ = common.points(k, seed=611)
    _, C, _, _ = common.points(n * k, seed=612, threads=8)
    B_, D0, _, _ = common.points(n, seed=613)
    c = C.reshape(n, k, 128)
    b, d0 = B_.reshape(n, 64), D0[:128]
    a = eprime
    got = schemes.bb04_ibe_decrypt_batch(engine, a, b, c, d0, dj.reshape(k, 64))
    for i in range(n):
        prod = None
        for j in range(k):
            t = port.pair_batch(dj[64 * j:64 * j + 64], c[i, j], 1)
            prod = t if prod is None else port.gt_mul_batch(prod, t, 1)
        mref = port.gt_div_batch(port.gt_mul_batch(a[i], prod, 1), port.pair_batch(b[i], d0, 1), 1)
        assert (got[i] == mref).all()
    # GWWW25: sum_j [coef_j] tauPowersG2_j
    Bn, n = 17, 3
    _, tau2, _, _ = common.points(Bn, seed=621, threads=8)
    coef = common.scalar_bytes(common.scalars(n * Bn, seed=622, edges=False)).reshape(n, Bn, 32)
    got = schemes.g2_msm_batch(engine, tau2.reshape(Bn, 128), coef)
    for i in range(n):
        acc = np.zeros(128, np.uint8)
        t = port.g2_mul_batch(tau2, coef[i].reshape(-1), Bn, 8)
        for j in range(Bn):
            acc = port.g2_add_batch(acc, t[128 * j:128 * j + 128], 1)
        assert (got[i] == acc).all()


def test_waters11_lw11_fused_drivers(engine):
    """§8f-3: Waters11 CP-ABE (cpabe/waters11/waters11_cpabe.go:248-290) and LW11 DABE (dabe/lw11_dabe.go:176-203,
    including its running-product exponent) -- fused batch drivers vs the reference's loops evaluated with the oracle."""
    from gopairingbasedcryptography_b200 import schemes

    n, m = 3, 4
    CX, DX, _, _ = common.points(n * m, seed=701, threads=8)
    KR, _, _, _ = common.points(m, seed=702)
    K1, CP, _, _ = common.points(n, seed=703)
    _, L, _, _ = common.points(1, seed=704)
    cx, dx = CX.reshape(n, m, 64), DX.reshape(n, m, 128)
    k, l = K1[:64], L[:128]
    cprime = CP.reshape(n, 128)
    krho = KR.reshape(m, 64)
    w = common.scalar_bytes(common.scalars(n * m, seed=705, edges=False)).reshape(n, m, 32)
    c = engine.pair_batch(K1, CP)  # any GT elements
    got = schemes.waters11_decrypt_batch(engine, c, cprime, cx, dx, k, l, krho, w)
    for i in range(n):
        den = None
        for j in range(m):
            t = port.gt_mul_batch(port.pair_batch(cx[i, j], l, 1), port.pair_batch(krho[j], dx[i, j], 1), 1)
            t = port.gt_exp_batch(t, w[i, j], 1)
            den = t if den is None else port.gt_mul_batch(den, t, 1)
        egs = port.gt_div_batch(port.pair_batch(k, cprime[i], 1), den, 1)
        assert (got[i] == port.gt_div_batch(c[i], egs, 1)).all()
    # LW11: D = 1; for x: D = (D * c1x * e(H, c3x) / e(Krho, c2x))^{w_x};  M = c0 / D
    C1P, C1Q, _, _ = common.points(n * m, seed=711, threads=8)
    c1x = engine.pair_batch(C1P, C1Q).reshape(n, m, 384)
    _, C2, _, _ = common.points(n * m, seed=712, threads=8)
    _, C3, _, _ = common.points(n * m, seed=713, threads=8)
    c2x, c3x = C2.reshape(n, m, 128), C3.reshape(n, m, 128)
    H, _, _, _ = common.points(1, seed=714)
    hgid = H[:64]
    got = schemes.lw11_decrypt_batch(engine, c, c1x, c2x, c3x, hgid, krho, w)
    one = np.frombuffer(o.gt_to_bytes(o.FP12_ONE), dtype=np.uint8)
    for i in range(n):
        D = one
        for x in range(m):
            D = port.gt_mul_batch(D, c1x[i, x], 1)
            D = port.gt_mul_batch(D, port.pair_batch(hgid, c3x[i, x], 1), 1)
            D = port.gt_div_batch(D, port.pair_batch(krho[x], c2x[i, x], 1), 1)
            D = port.gt_exp_batch(D, w[i, x], 1)
        assert (got[i] == port.gt_div_batch(c[i], D, 1)).all()


def test_small_batch_auto_routing(engine):
    """The default context sends pair / miller(k=1) / final-exp launches of <= 16384 elements to the lane-group kernels
    (5.8 ms instead of 10.6 ms for a small batch); both routes must give the same bytes as the oracle."""
    from gopairingbasedcryptography_b200 import bn254

    auto = bn254.default_engine()
    for n in (1, 33, 300):
        P, Q, _, _ = common.points(n, seed=900 + n, threads=8)
        if n > 3:
            P, Q = common.with_infinities(P, Q)
        ref = port.pair_batch(P, Q, n, 8)
        assert (auto.pair_batch(P, Q).reshape(-1) == ref).all()
        assert (engine.pair_batch(P, Q).reshape(-1) == ref).all()
        ml = auto.miller_loop_batch(P, Q, 1)
        assert (auto.final_exp_batch(ml).reshape(-1) == ref).all()
        assert (engine.final_exp_batch(ml).reshape(-1) == ref).all()
    assert (bn254.Pair([bn254.G1Affine(P[:64].tobytes())], [bn254.G2Affine(Q[:128].tobytes())]).raw ==
            port.pair_batch(P[:64], Q[:128], 1).tobytes())
    # products of 2..16 pairs with a small total pair count: Miller loop per pair on the lane-group kernels,
    # k-fold product, final exponentiation -- against the oracle and against the one-thread-per-product kernels
    for k, n in ((2, 100), (3, 64), (5, 40), (16, 7)):
        P, Q, _, _ = common.points(n * k, seed=950 + k, threads=8)
        P, Q = common.with_infinities(P, Q)
        ref = port.multi_pair_batch(P, Q, n, k, 8)
        assert (auto.multi_pair_batch(P, Q, k).reshape(-1) == ref).all(), k
        assert (auto.final_exp_batch(auto.miller_loop_batch(P, Q, k)).reshape(-1) == ref).all(), k
        okr = port.pairing_check_batch(P, Q, n, k, 8).astype(bool)
        assert (auto.pairing_check_batch(P, Q, k) == okr).all() and (engine.pairing_check_batch(P, Q, k) == okr).all()
    # a passing check through the routed path: e(A, B) e(-A, B) == 1
    A, B = P[:64].copy(), Q[:128].copy()
    nA = A.copy()
    nA[32:64] = np.frombuffer(((o.P - int.from_bytes(A[32:64].tobytes(), "little")) % o.P).to_bytes(32, "little"), dtype=np.uint8)
    assert auto.pairing_check_batch(np.concatenate([A, nA]), np.concatenate([B, B]), 2)[0]


def test_bsw07_round_trip_full_size(engine):
    """BASELINE configs[2] at its own size (batch 4096, 100-attribute AND policy) through a size-independent property:
    decrypt(encrypt(M)) == M for every ciphertext.  Keys and ciphertexts follow cpabe/bsw07/bsw07_cpabe.go:55-170 with its
    stub attribute hashes (H1 = g1, H2 = g2, bsw07_cpabe_utils.go:8-32); the 100 shares of each ciphertext are any values
    with sum_i Delta_i q(i) = s (equivalent to a random degree-99 polynomial with q(0) = s)."""
    from gopairingbasedcryptography_b200 import schemes

    n, m = 4096, 100
    R = o.R
    rng = o.SplitMix64(0xB2000254 + 3)
    g1, g2 = port.generators()
    alpha, beta, r = rng.scalar(), rng.scalar(), rng.scalar()
    rj = [rng.scalar() for _ in range(m)]
    inv_beta = pow(beta, -1, R)
    # user key (G2 side)
    d = engine.g2_mul_base_batch(g2, common.scalar_bytes([(alpha + r) * inv_beta % R]))[0]
    dj = engine.g2_mul_base_batch(g2, common.scalar_bytes([(r + x) % R for x in rj]))
    djp = engine.g2_mul_base_batch(g2, common.scalar_bytes(rj))
    # Lagrange coefficients at 0 for the points 1..m (access_tree_node.go:151-158)
    delta = []
    for i in range(1, m + 1):
        num, den = 1, 1
        for j in range(1, m + 1):
            if j != i:
                num = num * (-j) % R
                den = den * (i - j) % R
        delta.append(num * pow(den, -1, R) % R)
    # ciphertexts
    s = [rng.scalar() for _ in range(n)]
    msg = [rng.scalar() for _ in range(n)]
    inv_last = pow(delta[-1], -1, R)
    shares = np.empty((n, m, 32), dtype=np.uint8)
    for c in range(n):
        q = [rng.next() | (rng.next() << 64) | (rng.next() << 128) for _ in range(m - 1)]  # < 2^192 < r
        acc = sum(dl * qi for dl, qi in zip(delta, q)) % R
        q.append((s[c] - acc) * inv_last % R)
        shares[c] = np.frombuffer(b"".join(v.to_bytes(32, "little") for v in q), dtype=np.uint8).reshape(m, 32)
    cy = engine.g1_mul_base_batch(g1, shares.reshape(-1, 32)).reshape(n, m, 64)  # Cy = g1^q(i); Cy' = H1^q(i) = the same
    cc = engine.g1_mul_base_batch(g1, common.scalar_bytes([beta * x % R for x in s]))  # C = h^s, h = g1^beta
    e = port.pair_batch(g1, g2, 1)
    M = engine.gt_cyclo_exp_base_batch(e, common.scalar_bytes(msg))
    ctil = engine.gt_cyclo_exp_base_batch(e, common.scalar_bytes([(mm + alpha * x) % R for mm, x in zip(msg, s)]))
    db = common.scalar_bytes(delta).reshape(m, 32)
    pol = schemes.bsw07_policy_lines(engine, dj, djp, d, db)
    got = schemes.bsw07_decrypt_batch(engine, cy, cy, dj, djp, cc, d, ctil, db, lines=pol, folded=True)
    assert (got == M).all()
    assert (M[0] == port.gt_exp_batch(e, common.scalar_bytes([msg[0]]), 1)).all()
    # the other two decryption routes on a slice of the same batch
    k = 24
    key = schemes.bsw07_key_lines(engine, dj, djp, d)
    assert (schemes.bsw07_decrypt_batch(engine, cy[:k], cy[:k], dj, djp, cc[:k], d, ctil[:k], db, lines=key) == M[:k]).all()
    assert (schemes.bsw07_decrypt_batch(engine, cy[:k], cy[:k], dj, djp, cc[:k], d, ctil[:k], db) == M[:k]).all()


def test_setup_keygen_batches(engine):
    """§8f-4: tau-power ladders (afp25_bibe.go:156-162, gwww25_bibe.go:107-112), Waters05 SetUp (waters05_ibe.go:117-151)
    and BSW07 KeyGenerate (bsw07_cpabe.go:96-129) as fixed-base batches, against the oracle."""
    from gopairingbasedcryptography_b200 import schemes

    g1, g2 = port.generators()
    tau, B = 0x1234567890ABCDEF1234567890ABCDEF % o.R, 40
    pw = [pow(tau, i + 1, o.R) for i in range(B)]
    assert (schemes.tau_powers_g1(engine, tau, B).reshape(-1) == port.g1_mul_base_batch(g1, common.scalar_bytes(pw), B, 8)).all()
    assert (schemes.tau_powers_g2(engine, tau, B).reshape(-1) == port.g2_mul_base_batch(g2, common.scalar_bytes(pw), B, 8)).all()
    us = common.scalars(257, seed=81, edges=False)
    g1a, U = schemes.waters05_setup(engine, 987654321, us)
    assert (g1a == port.g1_mul_base_batch(g1, common.scalar_bytes([987654321]), 1)).all()
    assert (U.reshape(-1) == port.g2_mul_base_batch(g2, common.scalar_bytes(us), 257, 8)).all()
    alpha, beta, r = common.scalars(3, seed=82, edges=False)
    rj = common.scalars(7, seed=83, edges=False)
    g2a = port.g2_mul_base_batch(g2, common.scalar_bytes([alpha]), 1)
    d, dj, djp = schemes.bsw07_keygen(engine, g2a, beta, r, rj)
    assert (d == port.g2_mul_base_batch(g2, common.scalar_bytes([(alpha + r) * pow(beta, -1, o.R) % o.R]), 1)).all()
    assert (dj.reshape(-1) == port.g2_mul_base_batch(g2, common.scalar_bytes([(r + x) % o.R for x in rj]), 7, 8)).all()
    assert (djp.reshape(-1) == port.g2_mul_base_batch(g2, common.scalar_bytes(rj), 7, 8)).all()


def test_fixed_g1_pairing_check(engine):
    """bn254_pairing_check2_fixed_g1_batch (BLS verification shape) against the general check, on the thread kernels
    (fixture) and on the default context (small batches routed to the lane-group kernels), ragged size, with failures."""
    from gopairingbasedcryptography_b200 import bn254

    n = 300
    g1, g2 = port.generators()
    skb = common.scalar_bytes([0xC0FFEE1234567])
    pk = engine.g1_mul_base_batch(g1, skb)[0]
    hm = engine.g2_mul_base_batch(g2, common.scalar_bytes(common.scalars(n, seed=31, edges=False)))
    sig = engine.g2_mul_batch(hm, np.tile(skb, n))
    sig[7] = hm[7]          # wrong signature
    hm2 = hm.copy()
    hm2[11] = 0             # point at infinity in one pair -> that pair is skipped, check fails
    from gopairingbasedcryptography_b200 import schemes
    ng1 = schemes.neg_g1(g1)[0]
    P = np.concatenate([np.tile(pk.reshape(1, 64), (n, 1)), np.tile(ng1.reshape(1, 64), (n, 1))], axis=1)
    Qg = np.concatenate([hm2, sig], axis=1)
    ref = port.pairing_check_batch(P.reshape(-1), Qg.reshape(-1), n, 2, 8).astype(bool)
    assert not ref[7] and not ref[11] and ref.sum() == n - 2
    assert (engine.pairing_check2_fixed_g1_batch(pk, ng1, hm2, sig) == ref).all()
    assert (bn254.default_engine().pairing_check2_fixed_g1_batch(pk, ng1, hm2, sig) == ref).all()
    assert engine.pairing_check2_fixed_g1_batch(pk, ng1, b"", b"").shape == (0,)


def test_page_locked_caller_buffers(engine):
    """Caller buffers in page-locked memory take the direct-copy path of the host API (no staging memcpy); pageable and
    page-locked operands may be mixed and must give the same bytes."""
    import torch

    n = 3000
    P, Q, _, _ = common.points(n, seed=77, threads=8)
    ref = engine.pair_batch(P, Q)
    hP = torch.from_numpy(P.copy()).pin_memory().numpy()
    hQ = torch.from_numpy(Q.copy()).pin_memory().numpy()
    out = torch.empty((n, 384), dtype=torch.uint8).pin_memory().numpy()
    got = engine.pair_batch(hP, hQ, out=out)
    assert (got == ref).all() and (out == ref).all()
    assert (engine.pair_batch(hP, Q) == ref).all()           # mixed: pinned P, pageable Q, pageable result
    out2 = np.empty((n, 384), dtype=np.uint8)
    engine.pair_batch(P, hQ, out=out2)                          # pageable caller-owned result
    assert (out2 == ref).all()
    with pytest.raises(ValueError):
        engine.pair_batch(P, Q, out=np.empty((n - 1, 384), dtype=np.uint8))


def test_concurrent_callers(engine):
    """Host threads sharing one context serialise on its lock, and two contexts on one GPU run side by side (two Go
    goroutines calling into the library): every call returns the bytes of the single-threaded call."""
    import threading
    from gopairingbasedcryptography_b200 import bn254

    n = 2500
    P, Q, _, _ = common.points(n, seed=78, threads=8)
    sb = common.scalar_bytes(common.scalars(n, seed=79, edges=True))
    ref_pair = engine.pair_batch(P, Q)
    ref_mul = engine.g2_mul_batch(Q, sb)
    ref_chk = engine.pairing_check_batch(P, Q, 2)
    second = bn254.Engine(0)
    errs = []

    def worker(eng, kind):
        try:
            for _ in range(3):
                if kind == 0:
                    assert (eng.pair_batch(P, Q) == ref_pair).all()
                elif kind == 1:
                    assert (eng.g2_mul_batch(Q, sb) == ref_mul).all()
                else:
                    assert (eng.pairing_check_batch(P, Q, 2) == ref_chk).all()
        except Exception as e:  # noqa: BLE001
            errs.append(repr(e))

    ths = [threading.Thread(target=worker, args=(e, k)) for e in (engine, second) for k in (0, 1, 2)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    assert not errs, errs
