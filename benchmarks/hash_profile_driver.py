"""ncu driver: hash-to-G2 / hash-to-G1 over 2^17 32-byte messages (one launch each after a warm-up)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gopairingbasedcryptography_b200 import bn254  # noqa: E402

eng = bn254.default_engine()
n = 1 << 17
msgs = [i.to_bytes(32, "little") for i in range(n)]
for _ in range(2):
    g2 = eng.hash_to_g2_batch(msgs, b"BN254G2_XMD:SHA-256_SVDW_RO_")
    g1 = eng.hash_to_g1_batch(msgs, b"BN254G1_XMD:SHA-256_SVDW_RO_")
print(g1.shape, g2.shape)
