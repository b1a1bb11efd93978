import os, sys, time, json
sys.path.insert(0, "/root/repo")
import numpy as np
from gopairingbasedcryptography_b200 import bn254, schemes
e = bn254.Engine(0)
g1 = np.frombuffer(bn254.Generators()[2].raw, dtype=np.uint8).copy()
g2 = np.frombuffer(bn254.Generators()[3].raw, dtype=np.uint8).copy()
skb = bn254.scalars_to_bytes([123456789])
pk = e.g1_mul_base_batch(g1, skb)[0]
negg1 = schemes.neg_g1(g1.reshape(1, 64))[0]
for n in (6000, 8192, 10000, 12000, 16384, 20000, 24000):
    hm = e.g2_mul_base_batch(g2, bn254.scalars_to_bytes(list(range(7, 7 + n))))
    sig = e.g2_mul_batch(hm, np.tile(skb, (n, 1)))
    f = lambda: e.pairing_check2_fixed_g1_batch(pk, negg1, hm, sig)
    assert f().all()
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); f(); ts.append(time.perf_counter() - t0)
    print(json.dumps({"impl": os.environ.get("BN254_IMPL", "auto"), "n": n, "verify_ms": round(min(ts) * 1e3, 3)}), flush=True)
