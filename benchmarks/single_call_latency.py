#!/usr/bin/env python
"""What ONE 1-element call costs through the host API (the reference's own call shape: 1-element slices, one scalar, one
message) next to one thread of the C restatement of gnark -- the table a maintainer needs to decide where to batch.
Best of 7 per entry point, default routing.  One JSON line per row."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import common  # noqa: E402
from gopairingbasedcryptography_b200 import bn254, schemes  # noqa: E402
from oracle import port  # noqa: E402

e = bn254.Engine(0)
g1, g2 = port.generators()
sb = common.scalar_bytes(common.scalars(4, seed=77, edges=False))
P = e.g1_mul_base_batch(g1, sb)
Q = e.g2_mul_base_batch(g2, sb)
gt = e.pair_batch(P[:1], Q[:1])
negP = schemes.neg_g1(P[:1])
P2, Q2 = np.concatenate([P[:1], negP]), np.concatenate([Q[:1], Q[:1]])  # e(P, Q) e(-P, Q) = 1


def best(f, reps=7):
    f()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        f()
        ts.append(time.perf_counter() - t0)
    return min(ts) * 1e3


rows = [
    ("Pair (1 pair)", lambda: e.pair_batch(P[:1], Q[:1]), lambda: port.pair_batch(P[:1].reshape(-1), Q[:1].reshape(-1), 1)),
    ("PairingCheck (2 pairs)", lambda: e.pairing_check_batch(P2, Q2, 2), lambda: port.pairing_check_batch(P2.reshape(-1), Q2.reshape(-1), 1, 2)),
    ("G1 ScalarMultiplication", lambda: e.g1_mul_batch(P[:1], sb[:32]), lambda: port.g1_mul_batch(P[:1].reshape(-1), sb[:32], 1)),
    ("G2 ScalarMultiplication", lambda: e.g2_mul_batch(Q[:1], sb[:32]), lambda: port.g2_mul_batch(Q[:1].reshape(-1), sb[:32], 1)),
    ("G1 ScalarMultiplicationBase", lambda: e.g1_mul_base_batch(g1, sb[:32]), lambda: port.g1_mul_base_batch(g1, sb[:32], 1)),
    ("GT.Exp (GT proper)", lambda: e.gt_cyclo_exp_batch(gt, sb[:32]), lambda: port.gt_exp_batch(gt.reshape(-1), sb[:32], 1)),
    ("GT.Mul", lambda: e.gt_mul_batch(gt, gt), lambda: port.gt_mul_batch(gt.reshape(-1), gt.reshape(-1), 1)),
    ("HashToG1 (16-byte message)", lambda: e.hash_to_g1_batch([b"0123456789abcdef"], schemes.DST_BYTES_G1), None),
    ("HashToG2 (16-byte message)", lambda: e.hash_to_g2_batch([b"0123456789abcdef"], schemes.DST_BYTES_G2), None),
]
for name, gpu, cpu in rows:
    r = {"call": name, "b200_ms": round(best(gpu), 3)}
    if cpu is not None:
        r["cpu_restatement_1_thread_ms"] = round(best(cpu, reps=3), 3)
    print(json.dumps(r), flush=True)
e.close()
