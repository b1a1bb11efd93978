#!/usr/bin/env python
"""A/B of warp-VM INTERPRETER build variants (prefetch depth, LIN finish, LIN term batching) on the GPU box: one subprocess
per prebuilt variant library, device-resident Pair / MillerLoop / FinalExponentiation at 1 and 1024 items.
   python benchmarks/wvm_variants_ab.py "" wvm_ahead2 wvm_ahead3 ..."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TIMER = r'''
import os, sys, json, hashlib
sys.path.insert(0, %r)
import numpy as np, torch
os.environ["BN254_IMPL"] = "wvm"
from gopairingbasedcryptography_b200 import bn254
eng = bn254.Engine(0)
g1, g2 = bn254.Generators()[2:]
n = 1024
sb = bn254.scalars_to_bytes(list(range(3, 3 + n)))
P = eng.g1_mul_base_batch(g1.raw, sb); Q = eng.g2_mul_base_batch(g2.raw, sb)
dP, dQ = torch.from_numpy(P).cuda(), torch.from_numpy(Q).cuda(); dO = torch.empty((n, 384), dtype=torch.uint8, device="cuda"); dM = torch.empty_like(dO)
s = torch.cuda.current_stream().cuda_stream
def t(m, f):
    for _ in range(3): f(m)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): f(m)
    b.record(); torch.cuda.synchronize()
    return round(a.elapsed_time(b) / 10, 4)
pair = lambda m: eng.pair_batch_dev(dP.data_ptr(), dQ.data_ptr(), m, dO.data_ptr(), s)
ml = lambda m: eng.miller_loop_batch_dev(dP.data_ptr(), dQ.data_ptr(), m, 1, dM.data_ptr(), s)
fe = lambda m: eng.final_exp_batch_dev(dM.data_ptr(), m, dO.data_ptr(), s)
r = {"pair_ms_n1": t(1, pair), "miller_ms_n1": t(1, ml), "final_exp_ms_n1": t(1, fe), "pair_ms_n1024": t(1024, pair)}
pair(n); torch.cuda.synchronize()
r["sha"] = hashlib.sha256(dO.cpu().numpy().tobytes()).hexdigest()[:12]
print(json.dumps(r))
''' % ROOT
for v in sys.argv[1:]:
    env = dict(os.environ, BN254_VARIANT=v)
    out = subprocess.run([sys.executable, "-c", TIMER], env=env, capture_output=True, text=True)
    print(json.dumps({"variant": v or "default", **json.loads(out.stdout.strip().splitlines()[-1])}) if out.returncode == 0 else out.stderr[-600:], flush=True)
