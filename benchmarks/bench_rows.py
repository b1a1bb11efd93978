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
    "g1_var": 3.34e5, "g1_fixed": 4.9e4, "g2_var": 7.81e5, "gt_exp": 1.048e6, "bls_verify": 2.862e6, "bsw07_decrypt": 2.251e8,
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
    ng = 1 << (12 if args.quick else 15)
    gt = eng.pair_batch(Pn[:ng], Qn[:ng])
    sec, out = timed(lambda: eng.gt_exp_batch(gt, sb[:ng]), 2)
    cg = 256
    t0 = time.perf_counter(); ref = port.gt_exp_batch(gt[:cg].reshape(-1), sb[:cg], cg, T); cpu = cg / (time.perf_counter() - t0)
    assert (out[:cg].reshape(-1) == ref).all()
    row("gt_exp", "exps/s", ng, sec, cpu, "%d exps" % cg, "GT.Exp generic 254-bit square-and-multiply")
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
    nd = 64 if args.quick else 512
    cy = Pn[: nd * m].reshape(nd, m, 64)
    cyp = Pn[nd * m: 2 * nd * m].reshape(nd, m, 64)
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
    print(json.dumps({"summary": {r["row"]: r["value"] for r in rows}, "launches": eng.launches}))


if __name__ == "__main__":
    main()
