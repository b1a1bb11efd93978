"""Host-API time of the group kernels (GLV and fixed-base multiplications, hash-to-curve) at 2^17 elements, best of 4;
used to compare build variants (BN254_VARIANT=...)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common  # noqa: E402
from gopairingbasedcryptography_b200 import _build, bn254  # noqa: E402
from oracle import port  # noqa: E402

eng = bn254.default_engine()
n = 1 << 17
g1, g2 = port.generators()
sb = common.scalar_bytes(common.scalars(n, seed=5, edges=False))
P = eng.g1_mul_base_batch(g1, sb)
Q = eng.g2_mul_base_batch(g2, sb)
blob = np.frombuffer(b"".join(i.to_bytes(32, "little") for i in range(n)), dtype=np.uint8)
offs = np.arange(n + 1, dtype=np.uint64) * 32
rows = (("g1_glv", lambda: eng.g1_mul_batch(P, sb)), ("g2_glv", lambda: eng.g2_mul_batch(Q, sb)),
        ("g1_fixed", lambda: eng.g1_mul_base_batch(g1, sb)), ("g2_fixed", lambda: eng.g2_mul_base_batch(g2, sb)),
        ("hash_g1", lambda: eng.hash_to_g1_batch((blob, offs), b"D")), ("hash_g2", lambda: eng.hash_to_g2_batch((blob, offs), b"D")))
ref = {}
for name, fn in rows:
    best = 1e9
    for _ in range(4):
        t = time.perf_counter()
        out = fn()
        best = min(best, time.perf_counter() - t)
    print('{"variant": "%s", "row": "%s", "per_s": %.0f, "ms": %.3f, "crc": %d}'
          % (_build.VARIANT or "default", name, n / best, best * 1e3, int(np.frombuffer(out.tobytes(), dtype=np.uint32).sum(dtype=np.uint64))))
