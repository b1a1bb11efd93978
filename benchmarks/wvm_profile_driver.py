#!/usr/bin/env python
"""One 1-element Pair on the warp-VM latency kernels (for an ncu capture of a LONE warp: where do the ~1250 cycles of a
round go?).   ncu --set full --import-source on -k regex:k_wvm -c 1 -o out python benchmarks/wvm_profile_driver.py [n]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["BN254_IMPL"] = "wvm"
import numpy as np  # noqa: E402
import torch  # noqa: E402

from gopairingbasedcryptography_b200 import bn254  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
eng = bn254.Engine(0)
g1, g2 = bn254.Generators()[2:]
sb = bn254.scalars_to_bytes(list(range(3, 3 + n)))
P = eng.g1_mul_base_batch(g1.raw, sb)
Q = eng.g2_mul_base_batch(g2.raw, sb)
dP, dQ = torch.from_numpy(P).cuda(), torch.from_numpy(Q).cuda()
dO = torch.empty((n, 384), dtype=torch.uint8, device="cuda")
s = torch.cuda.current_stream().cuda_stream
eng.pair_batch_dev(dP.data_ptr(), dQ.data_ptr(), n, dO.data_ptr(), s)
torch.cuda.synchronize()
print("ok", int(dO.sum()))
eng.close()
