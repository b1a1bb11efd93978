#!/usr/bin/env python
"""A/B of the k_pair start stagger on the GPU box: one subprocess per setting (the library reads the environment once),
device-resident operands, CUDA events.   python benchmarks/stagger_ab.py [log2_n] "0 0" "1500 0" "4000 1" ...
Each argument is "<BN254_STAGGER cycles> <BN254_STAGGER_MODE>"."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TIMER = r'''
import os, sys, json
sys.path.insert(0, %r)
import numpy as np, torch
os.environ["BN254_IMPL"] = "thread"
from gopairingbasedcryptography_b200 import bn254
n = 1 << int(sys.argv[1])
eng = bn254.Engine(0)
g1, g2 = bn254.Generators()[2:]
sb = bn254.scalars_to_bytes(list(range(3, 3 + 4096)))
P = np.tile(eng.g1_mul_base_batch(g1.raw, sb), (n // 4096 + 1, 1))[:n]
Q = np.tile(eng.g2_mul_base_batch(g2.raw, sb), (n // 4096 + 1, 1))[:n]
dP, dQ = torch.from_numpy(P).cuda(), torch.from_numpy(Q).cuda(); dO = torch.empty((n, 384), dtype=torch.uint8, device="cuda")
s = torch.cuda.current_stream().cuda_stream
f = lambda: eng.pair_batch_dev(dP.data_ptr(), dQ.data_ptr(), n, dO.data_ptr(), s)
f(); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3): f()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 3
import hashlib
print(json.dumps({"ms": round(ms, 3), "pairings_per_s": round(n / ms * 1e3), "sha": hashlib.sha256(dO.cpu().numpy().tobytes()).hexdigest()[:16]}))
''' % ROOT

log2n = sys.argv[1]
for setting in sys.argv[2:]:
    st, mode = setting.split()
    env = dict(os.environ, BN254_STAGGER=st, BN254_STAGGER_MODE=mode)
    out = subprocess.run([sys.executable, "-c", TIMER, log2n], env=env, capture_output=True, text=True)
    print(json.dumps({"stagger": int(st), "mode": int(mode), **json.loads(out.stdout.strip().splitlines()[-1])}) if out.returncode == 0 else out.stderr[-800:], flush=True)
