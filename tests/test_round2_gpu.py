"""Round-2 GPU parity: explicit table handles, the fixed-base cache under concurrent callers, shared-point MSM,
device-resident entry points and pipelines, and the BASELINE.json configs at their stated sizes (configs[0] BLS with
the real hash at 1024 messages, configs[3] Waters05 encryption at 2^18, configs[4] AFP25 decryption at 1024 x 64).
Every comparison is bit-exact against the oracle (the reference's unfused formulas) or a size-independent property."""
import hashlib
import threading

import numpy as np
import pytest

from oracle import bn254_ref as o
from oracle import port

import common

pytestmark = pytest.mark.gpu
sb = common.scalar_bytes


def _torch_dev(engine):
    import torch

    return torch, torch.device("cuda", engine.device)


def test_fixed_base_handles(engine):
    """bn254_fixed_base_create + g1/g2_fixed_mul_batch + gt_fixed_exp_batch == the variable-base results (oracle)."""
    g1, g2 = port.generators()
    n = 300
    ks = common.scalars(n, seed=11, edges=True)
    base1 = port.g1_mul_base_batch(g1, sb([0xABCDEF123456789]), 1)
    base2 = port.g2_mul_base_batch(g2, sb([0x1234567]), 1)
    t1 = engine.fixed_base_create(1, base1)
    t2 = engine.fixed_base_create(2, base2)
    assert (engine.g1_fixed_mul_batch(t1, sb(ks)).reshape(-1) == port.g1_mul_base_batch(base1, sb(ks), n, 8)).all()
    assert (engine.g2_fixed_mul_batch(t2, sb(ks)).reshape(-1) == port.g2_mul_base_batch(base2, sb(ks), n, 8)).all()
    x = port.pair_batch(base1, base2, 1)
    t3 = engine.fixed_base_create(3, x)
    assert (engine.gt_fixed_exp_batch(t3, sb(ks[:64])).reshape(-1) == port.gt_exp_base_batch(x, sb(ks[:64]), 64, 8)).all()
    with pytest.raises(Exception):
        engine.g2_fixed_mul_batch(t1, sb(ks))  # wrong group
    # infinity base: every multiple is infinity
    t0 = engine.fixed_base_create(1, np.zeros(64, np.uint8))
    assert not engine.g1_fixed_mul_batch(t0, sb(ks[:8])).any()
    for t in (t0, t1, t2, t3):
        t.close()


def test_fixed_base_cache_two_threads_two_bases(engine):
    """The round-1 race (ADVICE: table rebuilt between lookup and launch): two host threads on ONE context alternate
    *_mul_base_batch calls on DIFFERENT bases, batches above the table threshold; every result must be the
    single-threaded one.  Also cycles through more bases than the cache holds."""
    g1, g2 = port.generators()
    n = 5000  # >= kFixedMin: table path
    ks = sb(common.scalars(n, seed=21, edges=True))
    bases1 = [port.g1_mul_base_batch(g1, sb([7 + 13 * i]), 1) for i in range(6)]
    bases2 = [port.g2_mul_base_batch(g2, sb([5 + 11 * i]), 1) for i in range(3)]
    ref1 = [port.g1_mul_base_batch(b, ks, n, 8) for b in bases1]
    ref2 = [port.g2_mul_base_batch(b, ks, n, 8) for b in bases2]
    x = port.pair_batch(bases1[0], bases2[0], 1)
    refx = port.gt_exp_base_batch(x, ks[:32 * 64], 64, 8)
    errs = []

    def worker(kind):
        try:
            for rep in range(3):
                if kind == 0:
                    for b, r in zip(bases1, ref1):
                        assert (engine.g1_mul_base_batch(b, ks).reshape(-1) == r).all(), "G1 base table mixed up"
                elif kind == 1:
                    for b, r in zip(reversed(bases1), reversed(ref1)):
                        assert (engine.g1_mul_base_batch(b, ks).reshape(-1) == r).all(), "G1 base table mixed up"
                else:
                    for b, r in zip(bases2, ref2):
                        assert (engine.g2_mul_base_batch(b, ks).reshape(-1) == r).all(), "G2 base table mixed up"
        except Exception as e:  # noqa: BLE001
            errs.append(repr(e))

    ths = [threading.Thread(target=worker, args=(k,)) for k in (0, 1, 2)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    assert not errs, errs
    # the GT cache entry shares the LRU with the group tables
    big = np.tile(ks[:32 * 64], 80)
    got = engine.gt_exp_base_batch(x, big).reshape(80, -1)
    assert (got[0] == refx).all() and (got[79] == refx).all()


@pytest.mark.parametrize("group,length,nvec", [(1, 1, 3), (1, 37, 5), (1, 256, 9), (2, 33, 4)])
def test_shared_point_msm_vs_oracle(engine, group, length, nvec):
    """bn254_msm_batch == the reference's loop of ScalarMultiplication + Add (afp25_bibe_utils.go:45-55), including
    zero scalars, r - 1, an infinity point and a vector of zeros."""
    P, Q, _, _ = common.points(length, seed=300 + length, threads=8)
    item = 64 if group == 1 else 128
    pts = (P if group == 1 else Q).copy().reshape(length, item)
    if length > 4:
        pts[3] = 0  # point at infinity among the bases
    ks = common.scalars(nvec * length, seed=301, edges=True)
    sc = sb(ks).reshape(nvec, length, 32).copy()
    sc[nvec - 1] = 0  # all-zero vector -> infinity
    table = engine.msm_table_create(group, pts)
    got = engine.msm_batch(table, sc)
    mul, add = (port.g1_mul_batch, port.g1_add_batch) if group == 1 else (port.g2_mul_batch, port.g2_add_batch)
    for v in range(nvec):
        terms = mul(pts.reshape(-1), sc[v].reshape(-1), length, 8).reshape(length, item)
        acc = np.zeros(item, np.uint8)
        for j in range(length):
            acc = add(acc, terms[j], 1)
        assert (got[v] == acc).all(), (group, length, v)
    assert not got[nvec - 1].any()
    table.close()


def test_device_resident_entry_points_match_host_paths(engine):
    """Every *_dev entry point added in round 2 against its host-buffer twin on the same inputs."""
    torch, dev = _torch_dev(engine)
    n = 200
    P, Q, _, _ = common.points(n, seed=41, threads=8)
    ks = sb(common.scalars(n, seed=42, edges=True))
    up = lambda a: torch.from_numpy(np.array(a, copy=True).reshape(-1)).to(dev)
    dn = lambda t: t.cpu().numpy()
    s = torch.cuda.current_stream().cuda_stream
    dP, dQ, dk = up(P), up(Q), up(ks)
    gt = engine.pair_batch(P, Q)
    dgt = up(gt)
    out_gt = torch.empty(n * 384, dtype=torch.uint8, device=dev)
    # GT mul / div, array x array and broadcast x array
    engine.dev("gt_mul_batch_dev", dgt.data_ptr(), 1, dgt.data_ptr(), 1, n, out_gt.data_ptr(), stream=s)
    assert (dn(out_gt).reshape(n, 384) == engine.gt_mul_batch(gt, gt)).all()
    engine.dev("gt_div_batch_dev", dgt.data_ptr(), 0, dgt.data_ptr(), 1, n, out_gt.data_ptr(), stream=s)
    assert (dn(out_gt).reshape(n, 384) == engine.gt_div_batch(np.tile(gt[0], (n, 1)), gt)).all()
    engine.dev("gt_cyclo_exp_batch_dev", dgt.data_ptr(), 1, dk.data_ptr(), n, out_gt.data_ptr(), stream=s)
    assert (dn(out_gt).reshape(n, 384) == engine.gt_cyclo_exp_batch(gt, ks)).all()
    engine.dev("gt_exp_batch_dev", dgt.data_ptr(), 1, dk.data_ptr(), n, out_gt.data_ptr(), stream=s)
    assert (dn(out_gt).reshape(n, 384) == engine.gt_exp_batch(gt, ks)).all()
    # group add / neg / sums
    o1 = torch.empty(n * 64, dtype=torch.uint8, device=dev)
    o2 = torch.empty(n * 128, dtype=torch.uint8, device=dev)
    P2, Q2 = np.roll(P, 64), np.roll(Q, 128)
    dP2, dQ2 = up(P2), up(Q2)  # keep every device operand referenced until its launch has been enqueued
    engine.dev("g1_add_batch_dev", dP.data_ptr(), dP2.data_ptr(), n, o1.data_ptr(), stream=s)
    assert (dn(o1).reshape(n, 64) == engine.g1_add_batch(P, P2)).all()
    engine.dev("g2_add_batch_dev", dQ.data_ptr(), dQ2.data_ptr(), n, o2.data_ptr(), stream=s)
    assert (dn(o2).reshape(n, 128) == engine.g2_add_batch(Q, Q2)).all()
    from gopairingbasedcryptography_b200 import schemes

    engine.dev("g1_neg_batch_dev", dP.data_ptr(), n, o1.data_ptr(), stream=s)
    assert (dn(o1).reshape(n, 64) == schemes.neg_g1(P)).all()
    engine.dev("g2_neg_batch_dev", dQ.data_ptr(), n, o2.data_ptr(), stream=s)
    assert (dn(o2).reshape(n, 128) == schemes.neg_g2(Q)).all()
    groups, length = 4, 50
    so = torch.empty(groups * 64, dtype=torch.uint8, device=dev)
    engine.dev("g1_sum_batch_dev", dP.data_ptr(), groups, length, so.data_ptr(), stream=s)
    assert (dn(so).reshape(groups, 64) == engine.g1_sum_batch(P[: groups * length * 64], length)).all()
    so2 = torch.empty(64, dtype=torch.uint8, device=dev)
    engine.dev("g1_sum_batch_dev", dP.data_ptr(), 1, n, so2.data_ptr(), stream=s)  # two passes (200 > 32)
    assert (dn(so2).reshape(1, 64) == engine.g1_sum_batch(P, n)).all()
    # subset sum (Waters hash)
    m = 64
    U = Q.reshape(n, 128)[: m + 1]
    sel = np.frombuffer(hashlib.sha256(b"sel").digest() * 8, dtype=np.uint8)[: 20 * (m // 8)].reshape(20, m // 8)
    ho = torch.empty(20 * 128, dtype=torch.uint8, device=dev)
    dU, dsel = up(U), up(sel)
    engine.dev("g2_subset_sum_batch_dev", dU.data_ptr(), m, dsel.data_ptr(), 20, ho.data_ptr(), stream=s)
    assert (dn(ho).reshape(20, 128) == engine.g2_subset_sum_batch(U, sel)).all()
    # BLS-shaped check, hash-to-curve, line tables
    ok = torch.empty(n, dtype=torch.uint8, device=dev)
    engine.dev("pairing_check2_fixed_g1_batch_dev", dP.data_ptr(), dQ.data_ptr(), dQ2.data_ptr(), n, ok.data_ptr(), stream=s)
    assert (dn(ok).astype(bool) == engine.pairing_check2_fixed_g1_batch(P[:64], P[64:128], Q, Q2)).all()
    msgs = [b"m%d" % i * (i % 5) for i in range(40)]
    blob = np.frombuffer(b"".join(msgs) or b"\0", dtype=np.uint8)
    off = np.zeros(41, dtype=np.uint64)
    off[1:] = np.cumsum([len(x) for x in msgs])
    dst = b"QUUX-V01-CS02-with-BN254G2_XMD:SHA-256_SVDW_RO_"
    hq = torch.empty(40 * 128, dtype=torch.uint8, device=dev)
    dblob, doff = up(blob), up(off.view(np.uint8))
    engine.dev("hash_to_g2_batch_dev", dblob.data_ptr(), doff.data_ptr(), 40, dst, hq.data_ptr(), stream=s)
    assert (dn(hq).reshape(40, 128) == engine.hash_to_g2_batch(msgs, dst)).all()
    hp = torch.empty(40 * 64, dtype=torch.uint8, device=dev)
    engine.dev("hash_to_g1_batch_dev", dblob.data_ptr(), doff.data_ptr(), 40, dst, hp.data_ptr(), stream=s)
    assert (dn(hp).reshape(40, 64) == engine.hash_to_g1_batch(msgs, dst)).all()
    m = 19
    lines = engine.g2_lines_create(Q.reshape(n, 128)[:m])
    rows = 10
    lo = torch.empty(rows * 384, dtype=torch.uint8, device=dev)
    engine.dev("multi_pair_lines_batch_dev", dP.data_ptr(), lines, rows, lo.data_ptr(), stream=s)
    assert (dn(lo).reshape(rows, 384) == engine.multi_pair_lines_batch(P[: rows * m * 64], lines)).all()
    # split multi-pairing (k > 16) on two different streams at once: stream-ordered scratch must not be shared
    k = 20
    s2 = torch.cuda.Stream()
    r1 = torch.empty(5 * 384, dtype=torch.uint8, device=dev)
    r2 = torch.empty(5 * 384, dtype=torch.uint8, device=dev)
    dPb, dQb = up(P[64 * 100:]), up(Q[128 * 100:])
    torch.cuda.synchronize()
    for _ in range(3):
        engine.dev("multi_pair_batch_dev", dP.data_ptr(), dQ.data_ptr(), 5, k, r1.data_ptr(), stream=s)
        engine.dev("multi_pair_batch_dev", dPb.data_ptr(), dQb.data_ptr(), 5, k, r2.data_ptr(), stream=s2.cuda_stream)
    torch.cuda.synchronize()
    assert (dn(r1).reshape(5, 384) == engine.multi_pair_batch(P[: 64 * 100], Q[: 128 * 100], k)).all()
    assert (dn(r2).reshape(5, 384) == engine.multi_pair_batch(P[64 * 100:], Q[128 * 100:], k)).all()


def test_bls_config0_exactly_1024_messages_real_hash(engine):
    """BASELINE configs[0]: sign + verify 1024 random messages, H(m) = hash.BytesToG2 (bls_signature.go:60,73) on the
    GPU, verification as the fixed-G1 check; 64 signatures are corrupted and must be rejected; H(m) and sigma of
    sampled messages are bit-exact against the oracle (definitional hash + C restatement)."""
    from gopairingbasedcryptography_b200 import schemes
    from oracle import hash_to_curve_ref as h2c

    n = 1024
    rng = o.SplitMix64(common.SEED + 0)
    msgs = [b"".join(rng.next().to_bytes(8, "little") for _ in range(4)) for _ in range(n)]
    sk = rng.scalar()
    g1, g2 = port.generators()
    pk = engine.g1_mul_base_batch(g1, sb([sk]))[0]
    hm = schemes.bytes_to_g2_batch(engine, msgs)
    sig = engine.g2_mul_batch(hm, np.tile(sb([sk]), n))
    for i in (0, 511, 1023):
        assert hm[i].tobytes() == o.g2_to_bytes(h2c.hash_to_g2(msgs[i], schemes.DST_BYTES_G2))
        assert (sig[i] == port.g2_mul_batch(hm[i], sb([sk]), 1)).all()
    neg_g1 = schemes.neg_g1(g1.reshape(1, 64))[0]
    ok = engine.pairing_check2_fixed_g1_batch(pk, neg_g1, hm, sig)
    assert ok.all()
    bad = sig.copy()
    flip = np.arange(0, n, 16)
    bad[flip] = sig[(flip + 1) % n]  # another message's signature
    ok = engine.pairing_check2_fixed_g1_batch(pk, neg_g1, hm, bad)
    expect = np.ones(n, bool)
    expect[flip] = False
    assert (ok == expect).all()
    # the reference's own call shape: PairingCheck([pk, -g1], [H(m), sigma]) per message
    chk = port.pairing_check_batch(np.concatenate([pk, neg_g1]), np.concatenate([hm[5], sig[5]]), 1, 2)
    assert bool(chk[0])


def test_waters05_encrypt_config3_at_2p18(engine):
    """BASELINE configs[3]: Waters05 encryption, batch 2^18, device-resident pipeline.  Sampled ciphertexts bit-exact
    against the reference's unfused flow (waters05_ibe.go:206-244 evaluated with the oracle); the WHOLE batch through
    two size-independent properties: c2 = [t]g1 re-derived by the variable-base kernel, and decryption of every 256th
    ciphertext recovers its message."""
    from gopairingbasedcryptography_b200 import schemes

    n, m = 1 << 18, 256
    g1, g2 = port.generators()
    alpha = 0x1234567890ABCDEF1234567890ABCDEF % o.R
    g1a, U = schemes.waters05_setup(engine, alpha, common.scalars(m + 1, seed=505, edges=False))
    params = schemes.Waters05Params(engine, g1a, U)
    ids = np.frombuffer(b"".join(hashlib.sha256(b"id-%d" % i).digest() for i in range(n)), dtype=np.uint8).reshape(n, 32)
    rng = o.SplitMix64(common.SEED + 3)
    ts = np.frombuffer(b"".join(o.scalar_to_bytes(rng.scalar()) for _ in range(n)), dtype=np.uint8).reshape(n, 32).copy()
    ts[0] = 0
    ts[1] = np.frombuffer(o.scalar_to_bytes(o.R - 1), dtype=np.uint8)
    ms = np.frombuffer(b"".join(o.scalar_to_bytes(rng.scalar()) for _ in range(512)), dtype=np.uint8)
    msg512 = engine.gt_cyclo_exp_base_batch(params.e_const, ms)
    msgs = np.tile(msg512, (n // 512, 1))
    c1, c2, c3 = schemes.waters05_encrypt_batch_dev(params, ids, msgs, ts)
    assert c1.shape == (n, 384) and c2.shape == (n, 64) and c3.shape == (n, 128)
    # (a) sampled, bit-exact vs the reference's unfused flow on the oracle
    e_const = port.pair_batch(g1a, g2, 1)
    assert (e_const == params.e_const).all()
    for i in (0, 1, 2, n // 2 + 77, n - 1):
        t = ts[i].tobytes()
        et = port.gt_exp_base_batch(e_const, np.frombuffer(t, np.uint8), 1)
        assert (c1[i] == port.gt_mul_batch(et, msgs[i], 1)).all()
        assert (c2[i] == port.g1_mul_base_batch(g1, np.frombuffer(t, np.uint8), 1)).all()
        acc = U[0].copy()
        for j in range(m):  # the Add loop of waters05_ibe.go:227-233
            if (ids[i, j >> 3] >> (7 - (j & 7))) & 1:
                acc = port.g2_add_batch(acc, U[j + 1], 1)
        assert (c3[i] == port.g2_mul_batch(acc, np.frombuffer(t, np.uint8), 1)).all()
    # (b) whole batch: c2 against the GLV kernel (independent code path)
    assert (engine.g1_mul_batch(np.tile(g1, n), ts) == c2).all()
    # (c) decrypt every 256th ciphertext: d1 = [alpha]g2 + [r]H(id), d2 = [r]g1; M = c1 * e(d2, c3) / e(c2, d1)
    idx = np.arange(0, n, 256)
    k = len(idx)
    H = engine.g2_subset_sum_batch(U, ids[idx])
    rs = sb(common.scalars(k, seed=507, edges=False)).reshape(k, 32)
    d1 = engine.g2_add_batch(np.tile(engine.g2_mul_base_batch(g2, sb([alpha])), (k, 1)), engine.g2_mul_batch(H, rs))
    d2 = engine.g1_mul_base_batch(g1, rs)
    P = np.concatenate([d2.reshape(k, 1, 64), schemes.neg_g1(c2[idx]).reshape(k, 1, 64)], axis=1)
    Q = np.concatenate([c3[idx].reshape(k, 1, 128), d1.reshape(k, 1, 128)], axis=1)
    rec = engine.gt_mul_batch(c1[idx], engine.multi_pair_batch(P, Q, 2))
    assert (rec == msgs[idx]).all()


def test_afp25_decrypt_config4_1024_identities_64_ciphertexts(engine):
    """BASELINE configs[4]: AFP25 batch decryption, B = 1024 identities (10000 + 10 i, afp25_bibe_test.go:381), 64
    ciphertexts for 64 distinct identities.  The whole Decrypt flow (afp25_bibe.go:369-418) -- quotient polynomial, pi =
    [q(tau)]1 as an MSM over the tau powers, the 3-pair product, Div -- against the reference's UNFUSED flow on the
    oracle for three ciphertexts, and decrypt(encrypt(M)) == M for all 64."""
    from gopairingbasedcryptography_b200 import bn254, schemes
    from test_fr_feeders import poly_from_roots

    B, n = 1024, 64
    R = o.R
    rng = o.SplitMix64(common.SEED + 4)
    tau, msk = rng.scalar(), rng.scalar()
    g1, g2 = port.generators()
    ids_int = [10000 + 10 * i for i in range(B)]
    tau_pows = schemes.tau_powers_g1(engine, tau, B)  # [tau]1 .. [tau^B]1  (afp25_bibe.go:156-162)
    g2_tau = engine.g2_mul_base_batch(g2, sb([tau]))[0]
    g2_msk = engine.g2_mul_base_batch(g2, sb([msk]))[0]
    table, f = schemes.afp25_batch_setup(engine, tau_pows, bn254.fr_from_ints(ids_int))
    f_int = bn254.fr_to_ints(f)
    f_tau = sum(c * pow(tau, k, R) for k, c in enumerate(f_int)) % R
    # Digest: D = [f(tau)]1 (afp25_bibe.go:293-307) through the same MSM machinery: coefficients over (g1, tau^1..tau^B)
    digest_table = engine.msm_table_create(1, np.concatenate([g1.reshape(1, 64), tau_pows], axis=0))
    D = engine.msm_batch(digest_table, bn254.fr_to_scalars(f).reshape(1, B + 1, 32))[0]
    assert (D == port.g1_mul_base_batch(g1, sb([f_tau]), 1)).all()
    ht = schemes.bytes_to_g1_batch(engine, [b"batch-0"])[0]  # h(t) = hash.BytesToG1(t.T)  (afp25_bibe_utils.go:10-12)
    sk = engine.g1_mul_batch(engine.g1_add_batch(D, ht), sb([msk]))[0]  # ComputeKey (afp25_bibe.go:329-336)
    # Encrypt (afp25_bibe.go:204-269) for 64 distinct identities of the batch
    who = [(37 * i + 5) % B for i in range(n)]
    r1 = [rng.scalar() for _ in range(n)]
    r2 = [rng.scalar() for _ in range(n)]
    msgs = engine.gt_cyclo_exp_base_batch(port.pair_batch(g1, g2, 1), sb([rng.scalar() for _ in range(n)]))
    g2_id = engine.g2_mul_base_batch(g2, sb([ids_int[w] for w in who]))
    a01 = engine.g2_add_batch(g2_id, schemes.neg_g2(np.tile(g2_tau, (n, 1))))  # [id]2 - [tau]2
    c1 = np.empty((n, 3, 128), np.uint8)
    c1[:, 0] = engine.g2_add_batch(engine.g2_mul_base_batch(g2, sb(r1)), engine.g2_mul_base_batch(g2_msk, sb(r2)))
    c1[:, 1] = engine.g2_mul_batch(a01, sb(r1))
    c1[:, 2] = schemes.neg_g2(engine.g2_mul_base_batch(g2, sb(r2)))
    b1 = engine.gt_div_batch(bn254._gt_one_raw(), engine.pair_batch(ht, g2_msk)[0])  # Inverse(e(h(t), [msk]2))
    c2 = engine.gt_mul_batch(engine.gt_exp_base_batch(b1[0], sb(r2)), msgs)
    # ---- the path under test
    ids_fr = bn254.fr_from_ints([ids_int[w] for w in who])
    got = schemes.afp25_decrypt_batch(engine, table, f, ids_fr, c1, c2, D, sk)
    assert (got == msgs).all(), "decrypt(encrypt(M)) != M"
    # ---- unfused reference flow on the oracle for three ciphertexts
    pts = np.concatenate([g1.reshape(1, 64), tau_pows[: B - 1]], axis=0)
    for v in (0, 31, 63):
        w = who[v]
        q = poly_from_roots(ids_int[:w] + ids_int[w + 1:])  # computePolynomialCoeffs(rootsWithoutId)
        terms = port.g1_mul_batch(pts.reshape(-1), sb(q), B, 8).reshape(B, 64)
        pi = np.zeros(64, np.uint8)
        for j in range(B):  # computeG1PolynomialTau: Add chain
            pi = port.g1_add_batch(pi, terms[j], 1)
        p1 = port.pair_batch(D, c1[v, 0], 1)
        p2 = port.pair_batch(pi, c1[v, 1], 1)
        p3 = port.pair_batch(sk, c1[v, 2], 1)
        den = port.gt_mul_batch(port.gt_mul_batch(p1, p2, 1), p3, 1)
        assert (got[v] == port.gt_div_batch(c2[v], den, 1)).all()
    table.close()
    digest_table.close()


def test_bsw07_device_pipeline_matches_host_pipeline(engine):
    """The device-resident BSW07 driver (key line tables; plain and folded) == schemes.bsw07_decrypt_batch."""
    from gopairingbasedcryptography_b200 import schemes

    n, m = 40, 12
    P, Q, _, _ = common.points(2 * n * m + n + 2 * m + 1, seed=61, threads=8)
    P, Q = P.reshape(-1, 64), Q.reshape(-1, 128)
    cy, cyp, c = P[: n * m].reshape(n, m, 64), P[n * m: 2 * n * m].reshape(n, m, 64), P[2 * n * m: 2 * n * m + n]
    dj, djp, d = Q[:m], Q[m: 2 * m], Q[2 * m]
    ct = engine.pair_batch(P[:n], Q[:n])
    deltas = sb(common.scalars(m, seed=62, edges=True)).reshape(m, 32)
    lines = schemes.bsw07_key_lines(engine, dj, djp, d)
    want = schemes.bsw07_decrypt_batch(engine, cy, cyp, dj, djp, c, d, ct, deltas, lines=lines)
    got = schemes.bsw07_decrypt_batch_dev(engine, cy, cyp, lines, c, ct, deltas)
    assert (got == want).all()
    folded = schemes.bsw07_policy_lines(engine, dj, djp, d, deltas)
    got2 = schemes.bsw07_decrypt_batch_dev(engine, cy, cyp, folded, c, ct, None)
    assert (got2 == want).all()


def test_warp_vm_kernels(engine):
    """BN254_IMPL=wvm: one warp per item (the latency path), through the same C ABI: pair / miller / final-exp / small
    multi-pairings and checks, ragged sizes, infinity operands, and a batch larger than one pass of the persistent grid."""
    import os

    from gopairingbasedcryptography_b200 import bn254

    old = os.environ.get("BN254_IMPL")
    os.environ["BN254_IMPL"] = "wvm"
    try:
        eng = bn254.Engine(0)
    finally:
        if old is None:
            del os.environ["BN254_IMPL"]
        else:
            os.environ["BN254_IMPL"] = old
    for n in (1, 5, 4 * 37 + 3, 5000):
        P, Q, _, _ = common.points(n, seed=1200 + n, threads=8)
        if n > 3:
            P, Q = common.with_infinities(P, Q)
        ref = port.pair_batch(P, Q, n, 8)
        assert (eng.pair_batch(P, Q).reshape(-1) == ref).all(), n
        if n <= 200:
            ml = eng.miller_loop_batch(P, Q, 1)
            assert (eng.final_exp_batch(ml).reshape(-1) == ref).all(), n
    rng = o.SplitMix64(13)
    x = np.frombuffer(b"".join(o.fp_to_mont_bytes(rng.fp()) for _ in range(36)), dtype=np.uint8).copy()
    assert (eng.final_exp_batch(x).reshape(-1) == port.final_exp_batch(x, 3)).all()
    zero_and_one = np.zeros(768, np.uint8)
    zero_and_one[384:416] = np.frombuffer(o.fp_to_mont_bytes(1), dtype=np.uint8)
    assert (eng.final_exp_batch(zero_and_one).reshape(-1) == port.final_exp_batch(zero_and_one, 2)).all()  # FE(0) = 0, FE(1) = 1
    for k, n in ((2, 100), (3, 64), (7, 9)):
        P, Q, _, _ = common.points(n * k, seed=1250 + k, threads=8)
        P, Q = common.with_infinities(P, Q)
        ref = port.multi_pair_batch(P, Q, n, k, 8)
        assert (eng.multi_pair_batch(P, Q, k).reshape(-1) == ref).all(), k
        okr = port.pairing_check_batch(P, Q, n, k, 8).astype(bool)
        assert (eng.pairing_check_batch(P, Q, k) == okr).all()
    # BLS shape through the fixed-G1 check
    n = 50
    P, Q, _, _ = common.points(2 * n + 2, seed=1300, threads=8)
    P, Q = P.reshape(-1, 64), Q.reshape(-1, 128)
    got = eng.pairing_check2_fixed_g1_batch(P[0], P[1], Q[2:2 + n], Q[2 + n:2 + 2 * n])
    Pc = np.tile(np.concatenate([P[0], P[1]]), n)
    Qc = np.concatenate([Q[2:2 + n], Q[2 + n:2 + 2 * n]], axis=1).reshape(-1)
    assert (got == port.pairing_check_batch(Pc, Qc, n, 2, 8).astype(bool)).all()
    eng.close()


def test_subset_sum_byte_window_tables(engine):
    """From 2048 selectors on the Waters hash runs from byte-window tables (one mixed addition per selector byte): same
    points as the plain per-bit sum, for G1 and G2, m not a multiple of 8, an infinity among the points, all-zero and
    all-one selectors."""
    n = 3000
    P, Q, _, _ = common.points(70, seed=1400, threads=8)
    rng = np.random.default_rng(5)
    for m, pts, item, f in ((20, P, 64, engine.g1_subset_sum_batch), (64, Q, 128, engine.g2_subset_sum_batch)):
        U = pts.reshape(-1, item)[: m + 1].copy()
        U[4] = 0  # point at infinity
        row = (m + 7) // 8
        sel = rng.integers(0, 256, size=(n, row), dtype=np.uint8)
        sel[0] = 0
        sel[1] = 255
        tab = f(U, sel)            # n >= 2048: table path
        plain = np.concatenate([f(U, sel[i:i + 1000]) for i in range(0, n, 1000)])  # chunks below the threshold: per-bit path
        assert (tab == plain).all()
        assert (tab[0] == U[0]).all()


def test_handles_may_outlive_their_context():
    """A garbage-collected host destroys objects in no particular order: table / line handles destroyed AFTER their
    context must neither hang nor crash (round 2: an interpreter with module-scope handles never exited), and the
    context's destruction releases their device memory."""
    import subprocess
    import sys

    code = r"""
import sys
sys.path.insert(0, %r)
import numpy as np
from gopairingbasedcryptography_b200 import bn254
eng = bn254.Engine(0)
g1 = np.frombuffer(bn254.Generators()[2].raw, dtype=np.uint8).copy()
g2 = np.frombuffer(bn254.Generators()[3].raw, dtype=np.uint8).copy()
s = bn254.scalars_to_bytes(list(range(2, 66)))
t1 = eng.fixed_base_create(1, g1)
P = eng.g1_fixed_mul_batch(t1, s)
Q = eng.g2_mul_base_batch(g2, s)
lines = eng.g2_lines_create(Q[:5])
msm = eng.msm_table_create(1, P[:8])
eng.close()                       # context first ...
for h in (t1, lines, msm):        # ... handles afterwards
    h.close()
print("ok")
""" % __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-800:]


def test_gt_cyclo_div_equals_generic_div_on_gt_elements(engine):
    """bn254_gt_cyclo_div_batch (a * conj(b)) == bn254_gt_div_batch (a * b^-1) == the oracle's GT.Div when b is a pairing
    output, a product or a power of pairing outputs; host and device-pointer entry points."""
    n = 70
    P, Q, _, _ = common.points(n, seed=0x6D1)
    b = engine.pair_batch(P, Q)
    b2 = engine.gt_mul_batch(b, b[::-1].copy())                    # products of pairing outputs
    b3 = engine.gt_cyclo_exp_batch(b, sb(common.scalars(n, seed=0x6D2)))   # powers (incl. the edge scalars 0, 1, r - 1 ...)
    a = engine.final_exp_batch(engine.miller_loop_batch(P[::-1].copy(), Q, 1))
    for den in (b, b2, b3):
        want = engine.gt_div_batch(a, den)
        assert (engine.gt_cyclo_div_batch(a, den) == want).all()
        assert (want.reshape(-1) == port.gt_div_batch(a.reshape(-1), den.reshape(-1), n)).all()
    torch, dev = _torch_dev(engine)
    da, db = torch.from_numpy(a).to(dev), torch.from_numpy(b2).to(dev)
    out = torch.empty_like(da)
    engine.dev("gt_cyclo_div_batch_dev", da.data_ptr(), 1, db.data_ptr(), 1, n, out.data_ptr(), stream=torch.cuda.current_stream(dev).cuda_stream)
    assert (out.cpu().numpy() == engine.gt_div_batch(a, b2)).all()


def test_two_pair_products_on_the_warp_vm(engine):
    """Small 2-pair products (BLS verification, every 2-pair PairingCheck) run ONE warp per product on the two-pair
    Miller program (k_wvm_miller2).  Against the oracle, including items whose first, second or both pairs contain the
    point at infinity (gnark skips such pairs: the kernel falls back to the single-pair program / to 1).  Batches below
    about 600 products stay on two warps per product (lower latency while schedulers are idle), hence n = 601."""
    n = 601
    P, Q, _, _ = common.points(2 * n, seed=0x2A1)
    P, Q = P.reshape(n, 2, 64).copy(), Q.reshape(n, 2, 128).copy()
    P[3, 0] = 0            # G1 infinity in the first pair
    Q[5, 1] = 0            # G2 infinity in the second pair
    P[7, 1] = 0; Q[7, 0] = 0   # both pairs skipped -> 1
    P[9, 0] = 0; P[9, 1] = 0
    ref = port.multi_pair_batch(P.reshape(-1), Q.reshape(-1), n, 2).reshape(n, 384)
    assert (engine.multi_pair_batch(P, Q, 2) == ref).all()
    assert (engine.multi_pair_batch(P[:41], Q[:41], 2) == ref[:41]).all()  # the two-warps-per-product route, same inputs
    ml = engine.miller_loop_batch(P, Q, 2)
    assert (engine.final_exp_batch(ml) == ref).all()
    ok = engine.pairing_check_batch(P, Q, 2)
    assert (ok == port.pairing_check_batch(P.reshape(-1), Q.reshape(-1), n, 2).astype(bool)).all()
    assert ok[7] and ok[9] and not ok[0]
    # the BLS shape: valid and forged signatures through the fixed-G1 entry point
    g1, g2 = port.generators()
    sk = sb([0x1234567])
    pk = engine.g1_mul_base_batch(g1, sk)[0]
    hm = Q[:, 0].copy(); hm[5] = Q[5, 0]
    sig = engine.g2_mul_batch(hm, np.tile(sk, n))
    sig[11] = sig[12]      # forged
    negg1 = port.g1_neg(g1) if hasattr(port, "g1_neg") else None
    from gopairingbasedcryptography_b200 import schemes

    negg1 = schemes.neg_g1(np.frombuffer(g1, dtype=np.uint8).reshape(1, 64))[0] if negg1 is None else negg1
    got = engine.pairing_check2_fixed_g1_batch(pk, negg1, hm, sig)
    Pc = np.concatenate([np.tile(pk.reshape(1, 64), (n, 1)), np.tile(np.asarray(negg1).reshape(1, 64), (n, 1))], axis=1).reshape(-1)
    Qc = np.concatenate([hm, sig], axis=1).reshape(-1)
    assert (got == port.pairing_check_batch(Pc, Qc, n, 2).astype(bool)).all()
    assert got.sum() == n - 1 and not got[11]


@pytest.mark.parametrize("group", [1, 2])
def test_fixed_base_four_items_per_thread_kernel(engine, group):
    """From 65 536 scalars on a fixed-base launch gives every thread four consecutive scalars and ONE inversion
    (Montgomery's trick): same affine points as the one-item kernel (two half-size calls), incl. the scalars 0, r, r + 1
    and 2^256 - 1, a batch size that is not a multiple of four, and sampled points against the oracle."""
    n = 70001
    g1, g2 = port.generators()
    base = np.frombuffer(g1 if group == 1 else g2, dtype=np.uint8)
    rng = np.random.default_rng(0xF1BED + group)
    s = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    edge = [0, o.R, o.R + 1, (1 << 256) - 1, 1, 2]
    for i, e in enumerate(edge):
        s[4 * i + (i % 4)] = np.frombuffer(int(e).to_bytes(32, "little"), dtype=np.uint8)
    s[n - 1] = 0
    table = engine.fixed_base_create(group, base)
    mul = engine.g1_fixed_mul_batch if group == 1 else engine.g2_fixed_mul_batch
    big = mul(table, s)
    half = np.concatenate([mul(table, s[:35000]), mul(table, s[35000:])], axis=0)
    assert (big == half).all()
    idx = np.concatenate([np.arange(24), [n - 3, n - 2, n - 1], rng.integers(0, n, 21)])
    ref_fn = port.g1_mul_base_batch if group == 1 else port.g2_mul_base_batch
    ref = ref_fn(base, s[idx].reshape(-1), len(idx), 4).reshape(len(idx), -1)
    assert (big[idx] == ref).all()
    table.close()


def test_group_kernels_with_shared_inversion_edge_cases(engine):
    """The group kernels share one inversion per CTA (every thread takes part: threads past the end and infinity
    operands run on substitutes).  Ragged batch sizes around the CTA size with infinity bases / operands in both groups,
    the zero scalar and r, against the oracle."""
    for n in (1, 127, 129, 257):
        P, Q, _, _ = common.points(n, seed=0x5EED + n)
        ks = common.scalars(n, seed=n)
        ks[0] = 0
        if n > 2:
            ks[2] = o.R
        k = sb(ks)
        P, Q = P.copy(), Q.copy()
        if n > 5:
            P[64 * 3:64 * 4] = 0
            Q[128 * 4:128 * 5] = 0
            Q[128 * (n - 1):] = 0
        assert (engine.g1_mul_batch(P, k).reshape(-1) == port.g1_mul_batch(P, k, n, 4)).all()
        assert (engine.g2_mul_batch(Q, k).reshape(-1) == port.g2_mul_batch(Q, k, n, 4)).all()
        Pr, Qr = np.roll(P, 64), np.roll(Q, 128)
        assert (engine.g1_add_batch(P, Pr).reshape(-1) == port.g1_add_batch(P, Pr, n, 4)).all()
        assert (engine.g2_add_batch(Q, Qr).reshape(-1) == port.g2_add_batch(Q, Qr, n, 4)).all()
        assert (engine.g2_add_batch(Q, Q).reshape(-1) == port.g2_add_batch(Q, Q, n, 4)).all()
        g1, g2 = port.generators()
        assert (engine.g2_mul_base_batch(g2, k).reshape(-1) == port.g2_mul_base_batch(g2, k, n, 4)).all()
        if n >= 16:
            m = 16
            sel = np.frombuffer(np.random.default_rng(n).bytes(n * 2), dtype=np.uint8).reshape(n, 2).copy()
            sel[0] = 0
            U = Q[:128 * (m + 1)].copy()
            got = engine.g2_subset_sum_batch(U, sel)
            for i in (0, 1, n // 2, n - 1):
                acc = U[:128].copy()
                for j in range(m):
                    if (sel[i, j >> 3] >> (7 - (j & 7))) & 1:
                        acc = port.g2_add_batch(acc, U[128 * (j + 1):128 * (j + 2)], 1, 1)
                assert (got[i] == acc).all()


def test_bb04_setup_keygen_encrypt_decrypt_at_reference_size(engine):
    """BB04-IBE at the reference's size (n = 256 identity bits, s = 2: the 512-point SetUp of ibe/bb04_ibe/bb04_ibe.go:109-136,
    KeyGenerate :138-168, Encrypt :170-206, Decrypt :210-241) through the batch drivers: sampled pieces bit-exact against
    the oracle's unfused formulas, and decrypt(encrypt(M)) == M for a batch of messages."""
    from gopairingbasedcryptography_b200 import schemes

    n, m = 256, 5
    g1, g2 = port.generators()
    alpha = common.scalars(1, seed=0xBB04, edges=False)[0]
    us = common.scalars(2 * n, seed=0xBB05, edges=False)
    g1a, u = schemes.bb04_setup(engine, alpha, us)
    assert u.shape == (n, 2, 128)
    assert g1a.tobytes() == port.g1_mul_base_batch(g1, sb([alpha]), 1, 1).tobytes()
    for idx in (0, 1, 255, 511):
        assert u.reshape(-1, 128)[idx].tobytes() == port.g2_mul_base_batch(g2, sb([us[idx]]), 1, 1).tobytes()
    ident = np.random.default_rng(4).integers(0, 2, n)
    g2a = port.g2_mul_base_batch(g2, sb([alpha]), 1, 1)
    rs = common.scalars(n, seed=0xBB06, edges=False)
    d0, dj = schemes.bb04_keygen(engine, g2a, u, ident, rs)
    assert (dj.reshape(-1) == port.g1_mul_base_batch(g1, sb(rs), n, 4)).all()
    acc = g2a.copy()   # the reference's chain: prod.Add(prod, [r_i] u[i][a_i]) then Add(g2^alpha, prod)
    for i in range(n):
        acc = port.g2_add_batch(acc, port.g2_mul_batch(u[i, ident[i]].copy().reshape(-1), sb([rs[i]]), 1, 1), 1, 1)
    assert d0.tobytes() == acc.tobytes()
    P, Q, _, _ = common.points(m, seed=0xBB07)
    msgs = engine.pair_batch(P, Q)
    ts = common.scalars(m, seed=0xBB08, edges=False)
    a, b, c = schemes.bb04_encrypt_batch(engine, g1a, u, ident, msgs, ts)
    k = port.gt_exp_batch(port.pair_batch(g1a, g2, 1), sb([ts[0]]), 1)
    assert a[0].tobytes() == port.gt_mul_batch(k, msgs[0].reshape(-1), 1).tobytes()
    assert b[0].tobytes() == port.g1_mul_base_batch(g1, sb([ts[0]]), 1, 1).tobytes()
    assert c[m - 1, 17].tobytes() == port.g2_mul_batch(u[17, ident[17]].copy().reshape(-1), sb([ts[m - 1]]), 1, 1).tobytes()
    back = schemes.bb04_ibe_decrypt_batch(engine, a, b, c, d0, dj)
    assert (back == msgs).all()


def test_gt_exp_four_dimensional_split_at_two_waves(engine):
    """GT.Exp on GT proper (four-dimensional split of the exponent, curve.cuh gt_cyclo_exp_gls4) over 2^17 elements:
    sampled outputs bit-exact against the oracle's ladder, the generic ladder and the size-independent group laws
    x^a x^b = x^(a + b mod r) and (x^a)^b = x^(a b mod r) on ALL of them, edge exponents included."""
    n = 1 << 17
    P, Q, _, _ = common.points(512, seed=0x6E1, threads=8)
    x = np.tile(engine.pair_batch(P, Q), (n // 512, 1))
    rng = o.SplitMix64(0x6E2)
    base = [rng.scalar() for _ in range(1021)]
    a = [base[i % 1021] ^ (i * 0x9E3779B97F4A7C15 & ((1 << 200) - 1)) for i in range(n)]
    b = [base[(7 * i + 3) % 1021] for i in range(n)]
    edges = [0, 1, 2, o.R - 1, o.R, o.R + 1, (1 << 256) - 1, 1 << 255, (1 << 128) - 1, 1 << 67, (1 << 67) - 1]
    a[:len(edges)] = edges
    b[n - len(edges):] = edges
    A, B = sb(a), sb(b)
    xa, xb = engine.gt_cyclo_exp_batch(x, A), engine.gt_cyclo_exp_batch(x, B)
    k = 24
    for lo in (0, n - k):  # oracle on the ends (the edge exponents live there)
        assert (xa[lo:lo + k].reshape(-1) == port.gt_exp_batch(x[lo:lo + k].reshape(-1), A[32 * lo:32 * (lo + k)], k)).all()
        assert (xb[lo:lo + k].reshape(-1) == port.gt_exp_batch(x[lo:lo + k].reshape(-1), B[32 * lo:32 * (lo + k)], k)).all()
    assert (engine.gt_exp_batch(x[:4096], A[:32 * 4096]) == xa[:4096]).all()          # generic ladder, same bytes
    s = sb([(u + v) % o.R for u, v in zip(a, b)])
    assert (engine.gt_mul_batch(xa, xb) == engine.gt_cyclo_exp_batch(x, s)).all()       # x^a x^b = x^(a + b)
    m = sb([(u * v) % o.R for u, v in zip(a, b)])
    assert (engine.gt_cyclo_exp_batch(xa, B) == engine.gt_cyclo_exp_batch(x, m)).all()  # (x^a)^b = x^(a b)
