"""Kernel-side time of the hash-to-curve kernels: host API call on 2^17 packed 32-byte messages, best of 5."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gopairingbasedcryptography_b200 import bn254  # noqa: E402

eng = bn254.default_engine()
n = 1 << 17
blob = np.frombuffer(b"".join(i.to_bytes(32, "little") for i in range(n)), dtype=np.uint8)
offs = np.arange(n + 1, dtype=np.uint64) * 32
for name, fn in (("g1", eng.hash_to_g1_batch), ("g2", eng.hash_to_g2_batch)):
    best = 1e9
    for _ in range(5):
        t = time.perf_counter()
        fn((blob, offs), b"BN254_XMD:SHA-256_SVDW_RO_")
        best = min(best, time.perf_counter() - t)
    print('{"row": "hash_to_%s_packed_2^17", "hashes_per_s": %.0f, "ms": %.3f}' % (name, n / best, best * 1e3))
