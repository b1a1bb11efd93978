#!/usr/bin/env python
"""Which module-scope object keeps benchmarks/profile_driver_r2.py from exiting?  Each case runs in its own interpreter
under a 60 s limit and leaves its objects alive at module scope (the interpreter's shutdown order destroys them)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PRE = r'''
import os, sys
sys.path.insert(0, %r)
os.environ["BN254_IMPL"] = "thread"
import numpy as np
from gopairingbasedcryptography_b200 import bn254, schemes
eng = bn254.Engine(0)
g1 = np.frombuffer(bn254.Generators()[2].raw, dtype=np.uint8).copy()
g2 = np.frombuffer(bn254.Generators()[3].raw, dtype=np.uint8).copy()
sb = bn254.scalars_to_bytes(list(range(5, 5 + 512)))
''' % ROOT
CASES = {
    "engine_only": "",
    "fixed_base": "t1 = eng.fixed_base_create(1, g1); P = eng.g1_fixed_mul_batch(t1, sb)",
    "fixed_base_g2_gt": "t2 = eng.fixed_base_create(2, g2); Q = eng.g2_fixed_mul_batch(t2, sb)\nP = eng.g1_mul_base_batch(g1, sb); gt = eng.pair_batch(P, Q); t3 = eng.fixed_base_create(3, gt[0]); eng.gt_fixed_exp_batch(t3, sb)",
    "key_lines": "P = eng.g1_mul_base_batch(g1, sb); Q = eng.g2_mul_base_batch(g2, sb)\nkey = schemes.bsw07_key_lines(eng, Q[:100], Q[100:200], Q[200])\neng.multi_pair_lines_batch(np.tile(P, (2, 1))[:201 * 4], key)",
    "msm_table": "P = eng.g1_mul_base_batch(g1, sb); table = eng.msm_table_create(1, P[:256]); eng.msm_batch(table, np.tile(sb, (1, 1))[:512])",
    "hash": "eng.hash_to_g2_batch([b'm %d' % i for i in range(3000)], schemes.DST_BYTES_G2)",
    "subset_sum": "Q = eng.g2_mul_base_batch(g2, sb); ids = np.frombuffer(np.random.default_rng(1).bytes(4096 * 32), dtype=np.uint8).reshape(4096, 32); eng.g2_subset_sum_batch(Q[:257], ids)",
    "three_engines": "P = eng.g1_mul_base_batch(g1, sb); Q = eng.g2_mul_base_batch(g2, sb)\nos.environ['BN254_IMPL'] = 'wvm'; ew = bn254.Engine(0); ew.pair_batch(P[:64], Q[:64])\nos.environ['BN254_IMPL'] = 'vm'; ev = bn254.Engine(0); ev.pair_batch(P[:64], Q[:64])",
}
for name, body in CASES.items():
    try:
        r = subprocess.run([sys.executable, "-X", "faulthandler", "-c", PRE + body + "\nprint('body done', flush=True)"], capture_output=True, text=True, timeout=90)
        print(name, "rc", r.returncode, r.stdout.strip()[-40:], r.stderr.strip()[-300:], flush=True)
    except subprocess.TimeoutExpired as e:
        print(name, "HANG at exit" if b"body done" in (e.stdout or b"") else "HANG in body", (e.stderr or b"")[-300:], flush=True)
