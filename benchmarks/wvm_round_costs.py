#!/usr/bin/env python
"""Cycles per warp-VM round by op class, from the BN254_VARIANT=wvmprof build (clock64 around every round of warp 0):
   BN254_VARIANT=wvmprof python benchmarks/wvm_round_costs.py"""
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["BN254_IMPL"] = "wvm"
from gopairingbasedcryptography_b200 import _build, _native, bn254  # noqa: E402

_build.build()
eng = bn254.Engine(0)
lib = _native.lib()
g1, g2 = bn254.Generators()[2:]
for n in (1, 592, 2368):
    sb = bn254.scalars_to_bytes(list(range(3, 3 + n)))
    P, Q = eng.g1_mul_batch(np.tile(np.frombuffer(g1.raw, np.uint8), n), sb), eng.g2_mul_batch(np.tile(np.frombuffer(g2.raw, np.uint8), n), sb)
    buf = (ctypes.c_ulonglong * 8)()
    lib.bn254_wvm_profile(buf)  # reset
    eng.pair_batch(P, Q)
    lib.bn254_wvm_profile(buf)
    v = list(buf)
    names = ["nop", "mul", "lin", "inv"]
    print(json.dumps({"n": n, **{names[i]: {"rounds": v[4 + i], "cycles_per_round": round(v[i] / v[4 + i], 1) if v[4 + i] else None} for i in range(1, 4)},
                      "total_cycles": sum(v[:4])}))
