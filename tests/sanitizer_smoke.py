"""Tiny run of every kernel family for compute-sanitizer (memcheck / racecheck), both pairing implementations.
   compute-sanitizer --tool memcheck python tests/sanitizer_smoke.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gopairingbasedcryptography_b200 import bn254  # noqa: E402
from oracle import port  # noqa: E402
import common  # noqa: E402

n = 11
P, Q, _, _ = common.points(2 * n, seed=1)
sb = common.scalar_bytes(common.scalars(2 * n))
eng = bn254.Engine(0)
ref = port.pair_batch(P[:64 * n], Q[:128 * n], n)
assert (eng.pair_batch(P[:64 * n], Q[:128 * n]).reshape(-1) == ref).all()
assert (eng.final_exp_batch(eng.miller_loop_batch(P[:64 * n], Q[:128 * n], 1)).reshape(-1) == ref).all()
assert (eng.multi_pair_batch(P, Q, 2).reshape(-1) == port.multi_pair_batch(P, Q, n, 2)).all()
assert (eng.multi_pair_batch(P[:64 * 20], Q[:128 * 20], 20).reshape(-1) == port.multi_pair_batch(P[:64 * 20], Q[:128 * 20], 1, 20)).all()
assert (eng.g1_mul_batch(P, sb).reshape(-1) == port.g1_mul_batch(P, sb, 2 * n)).all()
assert (eng.g2_mul_batch(Q, sb).reshape(-1) == port.g2_mul_batch(Q, sb, 2 * n)).all()
assert (eng.g1_add_batch(P, np.roll(P, 64)).reshape(-1) == port.g1_add_batch(P, np.roll(P, 64), 2 * n)).all()
gt = eng.pair_batch(P[:64 * 4], Q[:128 * 4])
assert (eng.gt_cyclo_exp_batch(gt, sb[:128]).reshape(-1) == port.gt_exp_batch(gt.reshape(-1), sb[:128], 4)).all()
assert (eng.gt_div_batch(gt, np.roll(gt, 1, axis=0)).reshape(-1) == port.gt_div_batch(gt.reshape(-1), np.roll(gt, 1, axis=0).reshape(-1), 4)).all()
sel = np.arange(32 * 3, dtype=np.uint8).reshape(3, 32)
eng.g2_subset_sum_batch(np.tile(Q[:128], 257), sel)
eng.g1_sum_batch(P, 11)
from oracle import bn254_ref as o  # noqa: E402
from oracle import hash_to_curve_ref as h2c  # noqa: E402
msgs = [b"", b"abc", b"x" * 150]
hm = eng.hash_to_g2_batch(msgs, h2c.DST_BYTES_G2)
assert hm[1].tobytes() == o.g2_to_bytes(h2c.hash_to_g2(b"abc", h2c.DST_BYTES_G2))
h1 = eng.hash_to_g1_batch(msgs, h2c.DST_BYTES_G1)
assert h1[2].tobytes() == o.g1_to_bytes(h2c.hash_to_g1(b"x" * 150, h2c.DST_BYTES_G1))
eng.pairing_check2_fixed_g1_batch(P[:64], P[64:128], Q[:128 * n], Q[128 * n:])
eng.pairing_check_batch(P, Q, 2)
print("sanitizer smoke ok, impl =", os.environ.get("BN254_IMPL", "thread"))
