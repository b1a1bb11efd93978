#!/usr/bin/env python
"""A/B of warp-VM program shapes ON the GPU box: regenerate the programs with different generator settings, rebuild the
library under a variant name, time a 1-element and a 192-element Pair.
   python benchmarks/wvm_ab.py "WVM_TINLINE=0" "WVM_TINLINE=15 WVM_DUP_MAX=6" ...   -> one JSON line per setting"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TIMER = r'''
import os, sys, json
sys.path.insert(0, %r)
import numpy as np, torch
os.environ["BN254_IMPL"] = "wvm"
from gopairingbasedcryptography_b200 import bn254
eng = bn254.Engine(0)
g1, g2 = bn254.Generators()[2:]
n = 192
sb = bn254.scalars_to_bytes(list(range(3, 3 + n)))
P = eng.g1_mul_batch(np.tile(np.frombuffer(g1.raw, np.uint8), n), sb); Q = eng.g2_mul_batch(np.tile(np.frombuffer(g2.raw, np.uint8), n), sb)
dP, dQ = torch.from_numpy(P).cuda(), torch.from_numpy(Q).cuda(); dO = torch.empty((n, 384), dtype=torch.uint8, device="cuda")
s = torch.cuda.current_stream().cuda_stream
def t(m, f):
    for _ in range(3): f(m)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): f(m)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / 10
pair = lambda m: eng.pair_batch_dev(dP.data_ptr(), dQ.data_ptr(), m, dO.data_ptr(), s)
ml = lambda m: eng.miller_loop_batch_dev(dP.data_ptr(), dQ.data_ptr(), m, 1, dO.data_ptr(), s)
print(json.dumps({"pair_ms_n1": round(t(1, pair), 4), "pair_ms_n192": round(t(192, pair), 4), "miller_ms_n1": round(t(1, ml), 4)}))
''' % ROOT

for i, setting in enumerate(sys.argv[1:]):
    env = dict(os.environ)
    for kv in setting.split():
        k, v = kv.split("=", 1)
        env[k] = v
    subprocess.check_call([sys.executable, os.path.join(ROOT, "gopairingbasedcryptography_b200", "csrc", "wvmgen.py")], env=env, stdout=subprocess.DEVNULL)
    env["BN254_VARIANT"] = "ab%d" % i
    env["BN254_KEEP_GENERATED"] = "1"
    subprocess.check_call([sys.executable, "-c", "import sys; sys.path.insert(0, %r); from gopairingbasedcryptography_b200 import _build; _build.VARIANTS[_build.VARIANT] = _build.DEFAULT; _build.build(force=True)" % ROOT], env=env)
    out = subprocess.run([sys.executable, "-c", "import sys; sys.path.insert(0, %r); from gopairingbasedcryptography_b200 import _build; _build.VARIANTS[_build.VARIANT] = _build.DEFAULT\n" % ROOT + TIMER], env=env, capture_output=True, text=True)
    print(json.dumps({"setting": setting, **json.loads(out.stdout.strip().splitlines()[-1])}) if out.returncode == 0 else out.stderr[-600:], flush=True)
