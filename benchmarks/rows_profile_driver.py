"""ncu driver: one launch of each non-headline kernel family at 56 832 elements (one full wave of 148 x 3 x 128 threads),
thread kernels (BN254_IMPL=thread), so a `--set full` capture gives the multiply-pipe activity of every SURVEY 8 row."""
import os
import sys

os.environ["BN254_IMPL"] = "thread"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import common  # noqa: E402
from gopairingbasedcryptography_b200 import bn254  # noqa: E402
from oracle import port  # noqa: E402

eng = bn254.Engine(0)
n = 148 * 3 * 128
g1, g2 = port.generators()
sb = common.scalar_bytes(common.scalars(n, seed=5, edges=False))
P = eng.g1_mul_base_batch(g1, sb)            # k_fixed_mul<G1> (+ table build)
Q = eng.g2_mul_base_batch(g2, sb)            # k_fixed_mul<G2>
eng.g1_mul_batch(P, sb)                      # k_scalar_mul<G1> (GLV)
eng.g2_mul_batch(Q, sb)                      # k_scalar_mul<G2>
ml = eng.miller_loop_batch(P, Q, 1)          # k_multi_pair_c<0,1>
gt = eng.final_exp_batch(ml)                 # k_final_exp
eng.gt_exp_batch(gt, sb)                     # k_gt_exp<0>
eng.gt_cyclo_exp_batch(gt, sb)               # k_gt_exp<1>
eng.gt_exp_base_batch(gt[0], sb)             # k_gt_fixed_exp
eng.gt_mul_batch(gt, gt)                     # k_gt_mul
eng.pairing_check2_fixed_g1_batch(P[0], P[1], Q, Q)   # k_check2_fixed_g1 (BLS verify shape)
eng.multi_pair_batch(np.tile(P[:n // 4], (4, 1)), np.tile(Q[:n // 4], (4, 1)), 4)  # k_multi_pair<0>, 4 pairs per product
print("done", eng.launches)
