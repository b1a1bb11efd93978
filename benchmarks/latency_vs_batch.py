#!/usr/bin/env python
"""Device-resident time of pair / miller(k=1) / final_exp for small batches: thread kernels, lane-group (K = 3 VM) kernels,
warp-VM kernels (one warp per item) and the default context's automatic routing.
   python benchmarks/latency_vs_batch.py  -> one JSON line per (impl, n); decides the small-batch routing thresholds."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gopairingbasedcryptography_b200 import bn254  # noqa: E402
from oracle import bn254_ref as o  # noqa: E402

engs = {}
os.environ["BN254_IMPL"] = "thread"
engs["thread"] = bn254.Engine(0)
os.environ["BN254_IMPL"] = "vm"
engs["vm"] = bn254.Engine(0)
os.environ["BN254_IMPL"] = "wvm"
engs["wvm"] = bn254.Engine(0)
os.environ.pop("BN254_IMPL", None)
engs["auto"] = bn254.Engine(0)
rng = o.SplitMix64(5)
N = 1 << 16
sb = bn254.scalars_to_bytes([rng.scalar() for _ in range(4096)])
g1, g2 = bn254.Generators()[2:]
e = engs["thread"]
P = np.tile(e.g1_mul_base_batch(g1.raw, sb), (N // 4096, 1))
Q = np.tile(e.g2_mul_base_batch(g2.raw, sb), (N // 4096, 1))
dP, dQ = torch.from_numpy(P).cuda(), torch.from_numpy(Q).cuda()
dO = torch.empty((N, 384), dtype=torch.uint8, device="cuda")
dM = torch.empty((N, 384), dtype=torch.uint8, device="cuda")
s = torch.cuda.current_stream().cuda_stream
engs["thread"].miller_loop_batch_dev(dP.data_ptr(), dQ.data_ptr(), N, 1, dM.data_ptr(), s)
torch.cuda.synchronize()


def t(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / 3


for n in (1, 64, 192, 1024, 2048, 3072, 4096, 6144, 8192, 16384, 32768, 65536):
    for name, eng in engs.items():
        if (name == "wvm" and n > 16384) or (name == "thread" and n < 64):
            continue
        r = {"impl": name, "n": n,
             "pair_ms": round(t(lambda: eng.pair_batch_dev(dP.data_ptr(), dQ.data_ptr(), n, dO.data_ptr(), s)), 3),
             "miller_ms": round(t(lambda: eng.miller_loop_batch_dev(dP.data_ptr(), dQ.data_ptr(), n, 1, dO.data_ptr(), s)), 3),
             "final_exp_ms": round(t(lambda: eng.final_exp_batch_dev(dM.data_ptr(), n, dO.data_ptr(), s)), 3)}
        print(json.dumps(r), flush=True)
