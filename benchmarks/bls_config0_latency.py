#!/usr/bin/env python
"""BASELINE configs[0] (bls01 sign + verify of 1024 messages) stage by stage through the host API, default routing:
hash.BytesToG2 / G2 ScalarMultiplication (sign) / the 2-pair fixed-G1 check (verify).  One JSON line."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from gopairingbasedcryptography_b200 import bn254, schemes  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
e = bn254.Engine(0)
g1 = np.frombuffer(bn254.Generators()[2].raw, dtype=np.uint8).copy()
skb = bn254.scalars_to_bytes([123456789])
pk = e.g1_mul_base_batch(g1, skb)[0]
negg1 = schemes.neg_g1(g1.reshape(1, 64))[0]
msgs = [b"bls01 message %08d" % i for i in range(n)]
sk_rows = np.tile(skb, (n, 1))


def best(f, reps=5):
    f()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        out = f()
        ts.append(time.perf_counter() - t0)
    return min(ts) * 1e3, out


t_hash, hm = best(lambda: schemes.bytes_to_g2_batch(e, msgs))
t_sign, sig = best(lambda: e.g2_mul_batch(hm, sk_rows))
t_ver, ok = best(lambda: e.pairing_check2_fixed_g1_batch(pk, negg1, hm, sig))
assert ok.all()
t_all, _ = best(lambda: e.pairing_check2_fixed_g1_batch(pk, negg1, *(lambda h: (h, e.g2_mul_batch(h, sk_rows)))(schemes.bytes_to_g2_batch(e, msgs))))
print(json.dumps({"messages": n, "hash_to_g2_ms": round(t_hash, 3), "sign_g2_mul_ms": round(t_sign, 3), "verify_check2_ms": round(t_ver, 3),
                  "hash_sign_verify_ms": round(t_all, 3), "messages_per_s": round(n / t_all * 1e3)}))
