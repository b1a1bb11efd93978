import sys, time
import os; ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import os
import numpy as np
from gopairingbasedcryptography_b200 import bn254, schemes
from oracle import port, bn254_ref as o
sys.path.insert(0, os.path.join(ROOT, 'tests')); import common
os.environ["BN254_IMPL"] = "thread"; thr = bn254.Engine(0); del os.environ["BN254_IMPL"]
auto = bn254.Engine(0)
n = 1024
g1, g2 = port.generators()
skb = common.scalar_bytes([123456789])
for name, e in (("thread", thr), ("auto", auto)):
    pk = e.g1_mul_base_batch(g1, skb)[0]
    msgs = [b"m%d" % i for i in range(n)]
    def flow():
        hm = schemes.bytes_to_g2_batch(e, msgs)
        sig = e.g2_mul_batch(hm, np.tile(skb, (n, 1)))
        return schemes.bls_verify_batch(e, pk, schemes.neg_g1(g1)[0], hm, sig), hm, sig
    flow()
    t0 = time.perf_counter(); ok, hm, sig = flow(); t1 = time.perf_counter()
    assert ok.all()
    t2 = time.perf_counter(); ok = schemes.bls_verify_batch(e, pk, schemes.neg_g1(g1)[0], hm, sig); t3 = time.perf_counter()
    print(name, "bls01 config0: 1024 msgs hash+sign+verify %.2f ms; verify only %.2f ms" % ((t1 - t0) * 1e3, (t3 - t2) * 1e3))
