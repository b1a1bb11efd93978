"""Host-API time of GT.Exp (generic and cyclotomic ladders) at several batch sizes, best of 3."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import common  # noqa: E402
from gopairingbasedcryptography_b200 import bn254  # noqa: E402
from oracle import port  # noqa: E402

eng = bn254.default_engine()
nmax = 1 << 17
g1, g2 = port.generators()
sb = common.scalar_bytes(common.scalars(nmax, seed=5, edges=False))
P = eng.g1_mul_base_batch(g1, sb[:32 * 4096])
Q = eng.g2_mul_base_batch(g2, sb[:32 * 4096])
gt = np.tile(eng.pair_batch(P, Q), (nmax // 4096, 1))
for n in (4096, 32768, 56832, 1 << 16, 1 << 17):
    for name, fn in (("gt_exp", eng.gt_exp_batch), ("gt_cyclo_exp", eng.gt_cyclo_exp_batch)):
        best = 1e9
        for _ in range(3):
            t = time.perf_counter()
            fn(gt[:n], sb[:32 * n])
            best = min(best, time.perf_counter() - t)
        print('{"row": "%s", "n": %d, "per_s": %.0f, "ms": %.3f}' % (name, n, n / best, best * 1e3))
