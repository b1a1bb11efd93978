#!/usr/bin/env python
"""A/B of the line-table Miller kernel's chunk size (table points per thread): one subprocess per prebuilt variant
library, the BSW07-100 shape (201 table points, 4096 products), device-resident.
   python benchmarks/lines_chunk_ab.py "" lines12 lines16"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TIMER = r'''
import os, sys, json, hashlib
sys.path.insert(0, %r)
import numpy as np, torch
from gopairingbasedcryptography_b200 import bn254
eng = bn254.Engine(0)
g1, g2 = bn254.Generators()[2:]
m, n = 201, 4096
sb = bn254.scalars_to_bytes(list(range(3, 3 + 4096)))
P = eng.g1_mul_base_batch(g1.raw, sb); Q = eng.g2_mul_base_batch(g2.raw, sb[:m])
lines = eng.g2_lines_create(Q)
Pd = np.tile(P, (n * m // 4096 + 1, 1))[: n * m]
dP = torch.from_numpy(Pd).cuda(); dO = torch.empty((n, 384), dtype=torch.uint8, device="cuda")
s = torch.cuda.current_stream().cuda_stream
f = lambda: eng.dev("multi_pair_lines_batch_dev", dP.data_ptr(), lines, n, dO.data_ptr(), stream=s)
f(); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3): f()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 3
print(json.dumps({"ms": round(ms, 3), "products_per_s": round(n / ms * 1e3), "sha": hashlib.sha256(dO.cpu().numpy().tobytes()).hexdigest()[:12]}))
lines.close(); eng.close()
''' % ROOT
for v in sys.argv[1:]:
    env = dict(os.environ, BN254_VARIANT=v)
    out = subprocess.run([sys.executable, "-c", TIMER], env=env, capture_output=True, text=True)
    print(json.dumps({"variant": v or "default (8)", **json.loads(out.stdout.strip().splitlines()[-1])}) if out.returncode == 0 else out.stderr[-600:], flush=True)
