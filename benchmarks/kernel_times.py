#!/usr/bin/env python
"""Kernel-only timings (device-resident operands, CUDA events) of the pairing-family entry points, to compare
their efficiency per Model-M Fp-mul:  python benchmarks/kernel_times.py [log2_n]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gopairingbasedcryptography_b200 import bn254  # noqa: E402
from oracle import bn254_ref as o  # noqa: E402

n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 17)
eng = bn254.default_engine()
rng = o.SplitMix64(99)
sb = bn254.scalars_to_bytes([rng.scalar() for _ in range(4096)])
g1, g2 = bn254.Generators()[2:]
P = np.tile(eng.g1_mul_base_batch(g1.raw, sb), (2 * n // 4096 + 1, 1))[:2 * n]
Q = np.tile(eng.g2_mul_base_batch(g2.raw, sb), (2 * n // 4096 + 1, 1))[:2 * n]
dP, dQ = torch.from_numpy(P).cuda(), torch.from_numpy(Q).cuda()
dO = torch.empty((n, 384), dtype=torch.uint8, device="cuda")
dB = torch.empty(n, dtype=torch.uint8, device="cuda")
s = torch.cuda.current_stream().cuda_stream
MODEL = {"pair": 15300, "miller_k1": 8049, "final_exp": 7251, "miller_k2": 2304 + 2 * 5743, "multi_pair_k2": 2304 + 2 * 5743 + 7251,
         "check_k2": 2304 + 2 * 5743 + 7251, "multi_pair_k8": 2304 + 8 * 5743 + 7251}


def t(name, fn, units):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(json.dumps({"kernel": name, "n": units, "ms": round(ms, 3), "per_s": units / ms * 1e3,
                      "G_fp_mul_per_s_modelM": units * MODEL[name] / ms / 1e6}))


t("pair", lambda: eng.pair_batch_dev(dP.data_ptr(), dQ.data_ptr(), n, dO.data_ptr(), s), n)
t("miller_k1", lambda: eng.miller_loop_batch_dev(dP.data_ptr(), dQ.data_ptr(), n, 1, dO.data_ptr(), s), n)
t("final_exp", lambda: eng.final_exp_batch_dev(dO.data_ptr(), n, dO.data_ptr(), s), n)
t("miller_k2", lambda: eng.miller_loop_batch_dev(dP.data_ptr(), dQ.data_ptr(), n, 2, dO.data_ptr(), s), n)
t("multi_pair_k2", lambda: eng.multi_pair_batch_dev(dP.data_ptr(), dQ.data_ptr(), n, 2, dO.data_ptr(), s), n)
t("check_k2", lambda: eng.pairing_check_batch_dev(dP.data_ptr(), dQ.data_ptr(), n, 2, dB.data_ptr(), s), n)
t("multi_pair_k8", lambda: eng.multi_pair_batch_dev(dP.data_ptr(), dQ.data_ptr(), n // 4, 8, dO.data_ptr(), s), n // 4)
