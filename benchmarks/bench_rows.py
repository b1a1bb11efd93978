#!/usr/bin/env python
"""Secondary measurements for the other SURVEY.md §8 rows and BASELINE.json configs (the headline line is
bench.py's).  Every number is taken through the public host-buffer API (H2D + kernels + D2H inside the timed
region) and printed beside the oracle's C restatement on the box's host threads (bounded samples).

  python benchmarks/bench_rows.py [--quick]  ->  one JSON object per row on stdout

Rows: G1/G2 variable-base (GLV) and fixed-base (window table) scalar mults/s, GT exp/s, BLS verifies/s
(config 1: 2-pair PairingCheck), BSW07 100-attribute fused decryptions/s (config 3), with the Model-M
limb-MAC roofline fraction of each (SURVEY.md §8d work figures x measured IMAD.WIDE peak).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

MAC_PEAK = 9.24e12  # measured IMAD.WIDE rate, profiles/microbench/imad_peak_b200.json
WORK = {  # SURVEY.md §8d limb-MACs per unit
    "g1_var": 3.34e5, "g1_fixed": 4.9e4, "g2_var": 7.81e5, "gt_exp": 1.048e6, "gt_cyclo_exp": 1.048e6, "bls_verify": 2.862e6,
    "bsw07_decrypt": 2.251e8, "waters05_encrypt": 1.577e6, "afp25_decrypt": 3.45e8,
}


def threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def timed(fn, reps):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    return (time.perf_counter() - t0) / reps, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    from gopairingbasedcryptography_b200 import bn254, schemes
    from oracle import bn254_ref as o
    from oracle import port

    eng = bn254.default_engine()
    T = threads()
    rng = o.SplitMix64(0xB2000254 + 7)
    g1, g2 = port.generators()
    n = 1 << (14 if args.quick else 18)

    def scal(m):
        return np.frombuffer(b"".join(o.scalar_to_bytes(rng.scalar()) for _ in range(m)), dtype=np.uint8).reshape(m, 32)

    sb = scal(4096)
    sb = np.tile(sb, (n // 4096 + 1, 1))[:n].copy()
    sb[:, 0] ^= np.arange(n, dtype=np.uint32).astype(np.uint8)  # distinct scalars
    rows = []

    def row(name, unit, n_units, sec, cpu_rate, cpu_sample, note=""):
        rate = n_units / sec
        rows.append({"row": name, "value": rate, "unit": unit, "n": n_units, "ms": sec * 1e3,
                     "roofline_frac_modelM": rate * WORK[name] / MAC_PEAK if name in WORK else None,
                     "cpu_baseline": {"value": cpu_rate, "cores": T, "kind": "port", "sample": cpu_sample},
                     "speedup_vs_cpu_port": rate / cpu_rate if cpu_rate else None, "note": note})
        print(json.dumps(rows[-1]), flush=True)

    # ---- bases ----------------------------------------------------------------------------------
    Pn = eng.g1_mul_base_batch(g1, sb)      # also builds the G1 table
    Qn = eng.g2_mul_base_batch(g2, sb)
    cs = 2048
    # G1 fixed base
    sec, _ = timed(lambda: eng.g1_mul_base_batch(g1, sb), 2)
    t0 = time.perf_counter(); ref = port.g1_mul_base_batch(g1, sb[:cs], cs, T); cpu = cs / (time.perf_counter() - t0)
    assert (Pn[:cs].reshape(-1) == ref).all()
    row("g1_fixed", "scalar-mults/s", n, sec, cpu, "%d mults" % cs, "ScalarMultiplicationBase: 32x255 window table")
    sec, _ = timed(lambda: eng.g2_mul_base_batch(g2, sb), 2)
    t0 = time.perf_counter(); ref = port.g2_mul_base_batch(g2, sb[:cs], cs, T); cpu = cs / (time.perf_counter() - t0)
    assert (Qn[:cs].reshape(-1) == ref).all()
    row("g2_fixed", "scalar-mults/s", n, sec, cpu, "%d mults" % cs, "G2 ScalarMultiplicationBase")
    # variable base (GLV)
    sb2 = np.roll(sb, 1, axis=0)
    sec, out = timed(lambda: eng.g1_mul_batch(Pn, sb2), 2)
    t0 = time.perf_counter(); ref = port.g1_mul_batch(Pn[:cs].reshape(-1), sb2[:cs], cs, T); cpu = cs / (time.perf_counter() - t0)
    assert (out[:cs].reshape(-1) == ref).all()
    row("g1_var", "scalar-mults/s", n, sec, cpu, "%d mults" % cs, "G1Affine.ScalarMultiplication, 2-dim GLV")
    sec, out = timed(lambda: eng.g2_mul_batch(Qn, sb2), 2)
    t0 = time.perf_counter(); ref = port.g2_mul_batch(Qn[:cs].reshape(-1), sb2[:cs], cs, T); cpu = cs / (time.perf_counter() - t0)
    assert (out[:cs].reshape(-1) == ref).all()
    row("g2_var", "scalar-mults/s", n, sec, cpu, "%d mults" % cs, "G2Affine.ScalarMultiplication (BLS sign), 2-dim GLV")
    # GT exp
    ng = 1 << (12 if args.quick else 17)
    gt = eng.pair_batch(Pn[:ng], Qn[:ng])
    sec, out = timed(lambda: eng.gt_exp_batch(gt, sb[:ng]), 2)
    cg = 256
    t0 = time.perf_counter(); ref = port.gt_exp_batch(gt[:cg].reshape(-1), sb[:cg], cg, T); cpu = cg / (time.perf_counter() - t0)
    assert (out[:cg].reshape(-1) == ref).all()
    row("gt_exp", "exps/s", ng, sec, cpu, "%d exps" % cg, "GT.Exp generic 254-bit square-and-multiply")
    sec, out2 = timed(lambda: eng.gt_cyclo_exp_batch(gt, sb[:ng]), 2)
    assert (out2[:cg].reshape(-1) == ref).all()
    row("gt_cyclo_exp", "exps/s", ng, sec, cpu, "%d exps (generic ladder, as gnark's GT.Exp)" % cg,
        "GT.Exp for pairing outputs: Granger-Scott squarings + signed window")
    # ---- config 1: BLS verify ---------------------------------------------------------------------
    nb = 1 << (12 if args.quick else 16)
    skb = sb[:1]
    pk = eng.g1_mul_base_batch(g1, skb)[0]
    hm = Qn[:nb]
    sigma = eng.g2_mul_batch(hm, np.tile(skb, (nb, 1)))
    neg = schemes.neg_g2(sigma)
    sec, ok = timed(lambda: schemes.bls_verify_batch(eng, pk, g1, hm, neg), 2)
    assert ok.all()
    cb = 1024
    Pc = np.concatenate([np.tile(pk.reshape(1, 64), (cb, 1)), np.tile(g1.reshape(1, 64), (cb, 1))], axis=1).reshape(-1)
    Qc = np.concatenate([hm[:cb], neg[:cb]], axis=1).reshape(-1)
    t0 = time.perf_counter(); okc = port.pairing_check_batch(Pc, Qc, cb, 2, T); cpu = cb / (time.perf_counter() - t0)
    assert okc.all()
    row("bls_verify", "verifies/s", nb, sec, cpu, "%d verifies" % cb,
        "config 1 (hash-to-G2 excluded: H(m) := [h]G2 synthetic hash, SURVEY 8d-1)")
    # ---- config 3: BSW07 100-attribute decrypt ---------------------------------------------------
    m = 100
    nd = 64 if args.quick else 2048
    big = np.tile(Pn, (2 * nd * m // n + 1, 1))
    cy = big[: nd * m].reshape(nd, m, 64)
    cyp = big[nd * m: 2 * nd * m].reshape(nd, m, 64)
    dj, djp = Qn[:m], Qn[m: 2 * m]
    c, d = Pn[-nd:], Qn[-1]
    ctil = gt[:nd] if nd <= ng else np.tile(gt, (nd // ng + 1, 1))[:nd]
    deltas = sb[:m]
    sec, out = timed(lambda: schemes.bsw07_decrypt_batch(eng, cy, cyp, dj, djp, c, d, ctil, deltas), 1)
    # CPU: the reference's unfused flow (2 Pair + Div + GT.Exp + Mul per leaf, then Pair + 2 Div), 2 decrypts per thread max
    cdn = max(2, min(T, 8))
    t0 = time.perf_counter()
    for i in range(cdn):
        e1 = port.pair_batch(cy[i].reshape(-1), dj.reshape(-1), m, T)
        e2 = port.pair_batch(cyp[i].reshape(-1), djp.reshape(-1), m, T)
        fz = port.gt_exp_batch(port.gt_div_batch(e1, e2, m, T), deltas.reshape(-1), m, T)
        A = fz[:384]
        for j in range(1, m):
            A = port.gt_mul_batch(A, fz[384 * j: 384 * j + 384], 1)
        ecd = port.pair_batch(c[i], d, 1)
        M = port.gt_div_batch(ctil[i], port.gt_div_batch(ecd, A, 1), 1)
        if i == 0:
            assert (out[0] == M).all()
    cpu = cdn / (time.perf_counter() - t0)
    row("bsw07_decrypt", "decryptions/s", nd, sec, cpu, "%d decryptions, unfused reference flow without its debug pairing" % cdn,
        "config 3: 200 G1 GLV mults + 201-pair Miller product + 1 final exp per decryption")
    key = schemes.bsw07_key_lines(eng, dj, djp, d)
    sec, out_l = timed(lambda: schemes.bsw07_decrypt_batch(eng, cy, cyp, dj, djp, c, d, ctil, deltas, lines=key), 1)
    assert (out_l == out).all()
    WORK["bsw07_decrypt_lines"] = WORK["bsw07_decrypt"]
    row("bsw07_decrypt_lines", "decryptions/s", nd, sec, cpu, "same CPU sample as bsw07_decrypt",
        "config 3 with the user key's 201 G2 points as precomputed line tables (no G2 arithmetic per ciphertext)")
    pol = schemes.bsw07_policy_lines(eng, dj, djp, d, deltas)
    nd4 = nd if args.quick else 4096  # BASELINE config 3: batch 4096
    reps4 = (nd4 + nd - 1) // nd
    cy4, cyp4 = np.tile(cy, (reps4, 1, 1))[:nd4], np.tile(cyp, (reps4, 1, 1))[:nd4]
    c4, ctil4 = np.tile(c, (reps4, 1))[:nd4], np.tile(ctil, (reps4, 1))[:nd4]
    sec, out_p = timed(lambda: schemes.bsw07_decrypt_batch(eng, cy4, cyp4, dj, djp, c4, d, ctil4, deltas, lines=pol, folded=True), 1)
    assert (out_p[:nd] == out).all()
    WORK["bsw07_decrypt_policy_lines"] = WORK["bsw07_decrypt"]
    row("bsw07_decrypt_policy_lines", "decryptions/s", nd4, sec, cpu, "same CPU sample as bsw07_decrypt",
        "config 3, batch %d: Lagrange coefficients folded into the key's line tables (e(C,[D]Q) = e(C,Q)^D): one 201-pair "
        "line-table product + 1 final exp per decryption, no per-ciphertext scalar multiplication" % nd4)
    # ---- config 4: Waters05 encrypt -------------------------------------------------------------------
    import hashlib
    nw = 1 << (12 if args.quick else 16)
    U = Qn[:257]
    e_const = gt[0]
    ids = np.stack([np.frombuffer(hashlib.sha256(b"id-%d" % i).digest(), dtype=np.uint8) for i in range(nw)])
    ts = sb[:nw]

    def waters_encrypt():
        c1 = eng.gt_mul_batch(eng.gt_cyclo_exp_base_batch(e_const, ts), gt[:nw] if nw <= ng else np.tile(gt, (nw // ng + 1, 1))[:nw])
        c2 = eng.g1_mul_base_batch(g1, ts)
        H = eng.g2_subset_sum_batch(U, ids)
        c3 = eng.g2_mul_batch(H, ts)
        return c1, c2, c3

    sec, (c1, c2, c3) = timed(waters_encrypt, 1)
    cw = 64
    t0 = time.perf_counter()
    e1 = port.gt_exp_base_batch(e_const, ts[:cw].reshape(-1), cw, T)
    r2 = port.g1_mul_base_batch(g1, ts[:cw].reshape(-1), cw, T)
    Hc = np.zeros((cw, 128), np.uint8)
    for i in range(cw):
        acc = U[0].copy()
        for j in range(256):
            if (ids[i, j >> 3] >> (7 - (j & 7))) & 1:
                acc = port.g2_add_batch(acc, U[j + 1], 1)
        Hc[i] = acc
    r3 = port.g2_mul_batch(Hc.reshape(-1), ts[:cw].reshape(-1), cw, T)
    cpu = cw / (time.perf_counter() - t0)
    assert (c2[:cw].reshape(-1) == r2).all() and (c3[:cw].reshape(-1) == r3).all()
    assert (c1[:cw].reshape(-1) == port.gt_mul_batch(e1, (gt[:cw]).reshape(-1), cw, T)).all()
    row("waters05_encrypt", "encryptions/s", nw, sec, cpu, "%d encryptions (Add loop single-threaded as in the reference)" % cw,
        "config 4: GT exp of the constant pairing + fixed-base G1 + Waters hash subset sum + G2 GLV mult")
    # ---- config 5: AFP25 decrypt shape (B-term G1 MSM + 3-pair product) ---------------------------------
    B = 1024
    nc = 8 if args.quick else 64
    tau = Pn[:B]
    coef = sb[: nc * B] if nc * B <= n else np.tile(sb, (nc * B // n + 1, 1))[: nc * B]

    def afp25():
        terms = eng.g1_mul_batch(np.tile(tau, (nc, 1)), coef)
        pi = eng.g1_sum_batch(terms, B)
        Pp = np.concatenate([pi.reshape(nc, 1, 64), Pn[B:B + nc].reshape(nc, 1, 64), Pn[2 * B:2 * B + nc].reshape(nc, 1, 64)], axis=1)
        Qp = np.tile(Qn[:3].reshape(1, 3, 128), (nc, 1, 1))
        return pi, eng.multi_pair_batch(Pp, Qp, 3)

    sec, (pi, prod) = timed(afp25, 1)
    t0 = time.perf_counter()
    tr = port.g1_mul_batch(tau.reshape(-1), coef[:B].reshape(-1), B, T)
    acc = np.zeros(64, np.uint8)
    for j in range(B):
        acc = port.g1_add_batch(acc, tr[64 * j:64 * j + 64], 1)
    cpu = 1 / (time.perf_counter() - t0)
    assert (pi[0] == acc).all()
    row("afp25_decrypt", "decryptions/s", nc, sec, cpu, "1 decryption's 1024-term MSM (mults on %d threads, Add chain serial as in the reference)" % T,
        "config 5 shape: 1024-term G1 MSM (GLV + segment sums) + 3-pair product per ciphertext; O(B^2) Fr polynomial stays on the host")
    # ---- §8f-1: hash-to-curve (BLS H(m), BF01 / LW11 identities): SHA-256 + SVDW + cofactor clearing on the GPU ----
    nh = 1 << (10 if args.quick else 16)
    msgs = [b"bls01 message %08d" % i for i in range(nh)]
    sec, hm = timed(lambda: schemes.bytes_to_g2_batch(eng, msgs), 1)
    row("hash_to_g2", "hashes/s", nh, sec, 0.0, "no CPU restatement in C (the Python oracle checks 4 outputs)",
        "hash.BytesToG2 = gnark HashToG2: expand_message_xmd(SHA-256), 2 SVDW maps over Fp2, add, psi cofactor clearing")
    blob = np.frombuffer(b"".join(msgs), dtype=np.uint8)
    offs = np.arange(nh + 1, dtype=np.uint64) * len(msgs[0])
    sec_p, hp = timed(lambda: eng.hash_to_g2_batch((blob, offs), schemes.DST_BYTES_G2), 1)
    assert (hp == hm).all()
    row("hash_to_g2_packed", "hashes/s", nh, sec_p, 0.0, "same outputs as hash_to_g2",
        "same, messages handed over as the C ABI takes them (one byte blob + offsets)")
    sec, h1 = timed(lambda: schemes.bytes_to_g1_batch(eng, msgs), 1)
    row("hash_to_g1", "hashes/s", nh, sec, 0.0, "no CPU restatement in C (the Python oracle checks 4 outputs)", "hash.BytesToG1 = gnark HashToG1")
    from oracle import hash_to_curve_ref as h2c
    for i in (0, 1, nh // 2, nh - 1):
        assert hm[i].tobytes() == o.g2_to_bytes(h2c.hash_to_g2(msgs[i], h2c.DST_BYTES_G2))
        assert h1[i].tobytes() == o.g1_to_bytes(h2c.hash_to_g1(msgs[i], h2c.DST_BYTES_G1))
    # config 1 end to end with the real hash: hash + sign (G2 GLV mult) + verify (2-pair check) per message
    skb = sb[:1]
    pk = eng.g1_mul_base_batch(g1, skb)[0]

    def bls_full():
        hmm = schemes.bytes_to_g2_batch(eng, msgs)
        sig = eng.g2_mul_batch(hmm, np.tile(skb, (nh, 1)))
        return schemes.bls_verify_batch(eng, pk, schemes.neg_g1(g1)[0], hmm, sig)

    sec, okf = timed(bls_full, 1)
    assert okf.all()
    row("bls_hash_sign_verify", "messages/s", nh, sec, 0.0, "-", "config 1 end to end: H(m) + sign + verify per message, all on the GPU")
    # BASELINE configs[0] at its own size: 1024 messages, hash + sign + verify.  A batch this small is latency-bound;
    # the default context runs its Miller loops / final exponentiations on the lane-group kernels.
    m0 = msgs[:1024]

    def bls_1024():
        hmm = schemes.bytes_to_g2_batch(eng, m0)
        sig = eng.g2_mul_batch(hmm, np.tile(skb, (1024, 1)))
        return schemes.bls_verify_batch(eng, pk, schemes.neg_g1(g1)[0], hmm, sig), hmm, sig

    sec, (ok0, hm0, sig0) = timed(bls_1024, 3)
    assert ok0.all()
    t0 = time.perf_counter()
    sgc = port.g2_mul_batch(hm0.reshape(-1), np.tile(skb, (1024, 1)).reshape(-1), 1024, T)
    Pc0 = np.concatenate([np.tile(pk.reshape(1, 64), (1024, 1)), np.tile(schemes.neg_g1(g1)[0].reshape(1, 64), (1024, 1))], axis=1).reshape(-1)
    Qc0 = np.concatenate([hm0, sgc.reshape(1024, 128)], axis=1).reshape(-1)
    assert port.pairing_check_batch(Pc0, Qc0, 1024, 2, T).all()
    cpu0 = 1024 / (time.perf_counter() - t0)
    row("bls01_config0_1024", "messages/s", 1024, sec, cpu0, "sign + verify of the same 1024 messages on %d threads (hash-to-G2 not included: no C restatement)" % T,
        "BASELINE configs[0]: hash + sign + verify of 1024 messages, %.1f ms per batch" % (sec * 1e3))
    print(json.dumps({"summary": {r["row"]: r["value"] for r in rows}, "launches": eng.launches}))


if __name__ == "__main__":
    main()
