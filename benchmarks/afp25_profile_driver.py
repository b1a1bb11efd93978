#!/usr/bin/env python
"""BASELINE configs[4] (AFP25 batch decryption, B = 1024 identities, 64 ciphertexts) as a stand-alone driver: prints
the host-API time per batch; under `ncu --metrics gpu__time_duration.sum` the launch list shows where it goes."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from benchmarks.rows import SplitMix64, scalar_block  # noqa: E402
from gopairingbasedcryptography_b200 import bn254, schemes  # noqa: E402

eng = bn254.Engine(0)
R = bn254.R_MOD
rng = SplitMix64(11)
g1 = np.frombuffer(bn254.Generators()[2].raw, dtype=np.uint8).copy()
g2 = np.frombuffer(bn254.Generators()[3].raw, dtype=np.uint8).copy()
B, nc = 1024, int(sys.argv[1]) if len(sys.argv) > 1 else 64
sb = scalar_block(rng, 256, R)
Pn, Qn = eng.g1_mul_base_batch(g1, sb), eng.g2_mul_base_batch(g2, sb)
gt = eng.pair_batch(Pn[:nc], Qn[:nc])
ids_int = [10000 + 10 * i for i in range(B)]
tau_pows = schemes.tau_powers_g1(eng, rng.scalar(R), B)
table, f = schemes.afp25_batch_setup(eng, tau_pows, bn254.fr_from_ints(ids_int))
who = [(37 * i + 5) % B for i in range(nc)]
ids_fr = bn254.fr_from_ints([ids_int[w] for w in who])
c1a = np.tile(Qn[:3].reshape(1, 3, 128), (nc, 1, 1))
run = lambda: schemes.afp25_decrypt_batch(eng, table, f, ids_fr, c1a, gt[:nc], Pn[7], Pn[9])
out = run()
torch.cuda.synchronize()
ts = []
for _ in range(5):
    t0 = time.perf_counter()
    out = run()
    ts.append(time.perf_counter() - t0)
print(json.dumps({"ciphertexts": nc, "identities": B, "ms_per_batch_best": round(min(ts) * 1e3, 3), "ms_per_batch_median": round(sorted(ts)[2] * 1e3, 3),
                  "decryptions_per_s": round(nc / min(ts))}), flush=True)
table.close()
eng.close()
