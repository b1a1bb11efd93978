"""ncu driver (round 2): ONE launch of every kernel family at one full wave (56 832 = 148 x 3 x 128 elements) on the thread
kernels, plus the warp-VM and lane-group kernels at their own grid size and the table-driven kernels at the shapes the
BASELINE configs use.  A `--metrics` pass over this script gives, per kernel: time, registers, occupancy, multiply-pipe
activity, executed fmaheavy (IMAD-class) instructions and DRAM bytes -> profiles/r2/ncu_kernel_families_*.
   python benchmarks/profile_driver_r2.py            (prints the unit count of every launch, in launch order)"""
import json
import os
import sys

os.environ["BN254_IMPL"] = "thread"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from benchmarks.rows import SplitMix64, scalar_block  # noqa: E402
from gopairingbasedcryptography_b200 import bn254, schemes  # noqa: E402

import faulthandler  # noqa: E402

faulthandler.dump_traceback_later(int(os.environ.get("BN254_DRIVER_WATCHDOG_S", "600")), exit=True)
T0 = __import__("time").time()
eng = bn254.Engine(0)
n = 148 * 3 * 128
R = bn254.R_MOD
rng = SplitMix64(5)
g1 = np.frombuffer(bn254.Generators()[2].raw, dtype=np.uint8).copy()
g2 = np.frombuffer(bn254.Generators()[3].raw, dtype=np.uint8).copy()
sb = scalar_block(rng, n, R)
units = []


def mark(kernel, count, what):
    print("%7.1f s  %s" % (__import__("time").time() - T0, kernel), file=sys.stderr, flush=True)
    units.append({"kernel": kernel, "units": count, "what": what})


t1, t2 = eng.fixed_base_create(1, g1), eng.fixed_base_create(2, g2)   # table builds: 8160-element k_scalar_mul launches
P = eng.g1_fixed_mul_batch(t1, sb); mark("k_fixed_mul<G1", n, "G1 fixed-base mults")
Q = eng.g2_fixed_mul_batch(t2, sb); mark("k_fixed_mul<G2", n, "G2 fixed-base mults")
eng.g1_mul_batch(P, sb); mark("k_scalar_mul<bn254::G1Jac", n, "G1 GLV mults")
eng.g2_mul_batch(Q, sb); mark("k_scalar_mul_g2_gls", n, "G2 GLS-4 mults")
ml = eng.miller_loop_batch(P, Q, 1); mark("k_multi_pair_c<0, 1>", n, "Miller loops")
gt = eng.final_exp_batch(ml); mark("k_final_exp", n, "final exponentiations")
eng.pair_batch(P, Q); mark("k_pair", n, "pairings")
eng.gt_exp_batch(gt, sb); mark("k_gt_exp<0>", n, "generic GT.Exp")
eng.gt_cyclo_exp_batch(gt, sb); mark("k_gt_exp<1>", n, "GT-proper GT.Exp")
t3 = eng.fixed_base_create(3, gt[0])
eng.gt_fixed_exp_batch(t3, sb); mark("k_gt_fixed_exp", n, "fixed-base GT.Exp")
eng.gt_mul_batch(gt, gt); mark("k_gt_mul<0>", n, "GT products")
eng.pairing_check2_fixed_g1_batch(P[0], P[1], Q, Q); mark("k_check2_fixed_g1", n, "BLS verifications")
# BSW07 shape: 201-pair products from line tables, 4 waves of (item, group) threads
m = 100
nd = 2048
key = schemes.bsw07_key_lines(eng, Q[:m], Q[m:2 * m], Q[2 * m])
Pd = np.tile(P, (nd * (2 * m + 1) // n + 1, 1))[: nd * (2 * m + 1)]
eng.multi_pair_lines_batch(Pd, key); mark("k_miller_lines", nd, "BSW07 decryptions (201-pair line-table products)")
# Waters hash, byte-window tables
ids = np.frombuffer(np.random.default_rng(1).bytes(n * 32), dtype=np.uint8).reshape(n, 32)
eng.g2_subset_sum_batch(Q[:257], ids); mark("k_subset_sum_tab<bn254::G2Jac", n, "Waters hashes (256-bit identities)")
# shared-point MSM, AFP25 shape
B, nv = 1024, 64
table = eng.msm_table_create(1, P[:B]); mark("k_msm_tables<bn254::G1Jac", B, "points (32 x 255-entry window tables each)")
eng.msm_batch(table, np.tile(sb, (B * nv // n + 1, 1))[: B * nv]); mark("k_msm_partial<bn254::G1Jac", nv, "1024-term MSMs")
# hash-to-curve
msgs = [b"bls01 message %08d" % i for i in range(n)]
eng.hash_to_g2_batch(msgs, schemes.DST_BYTES_G2); mark("k_hash_to_curve<2>", n, "HashToG2")
# latency kernels
os.environ["BN254_IMPL"] = "wvm"
ew = bn254.Engine(0)
ew.pair_batch(P[:2368], Q[:2368]); mark("k_wvm<1>", 2368, "pairings, one warp each (one pass of the grid)")
os.environ["BN254_IMPL"] = "vm"
ev = bn254.Engine(0)
ev.pair_batch(P[:17760], Q[:17760]); mark("k_vm<", 17760, "pairings, three lanes each (one pass of the grid)")
print(json.dumps(units), flush=True)
# explicit teardown, handles before their contexts (timestamps on stderr: a profiler attached to this process waits for it)
import time  # noqa: E402

for name in ("key", "table", "t3", "t2", "t1", "ev", "ew", "eng"):
    t0 = time.time()
    obj = globals().pop(name)
    if hasattr(obj, "close"):
        obj.close()
    del obj
    print("closed %s in %.2f s" % (name, time.time() - t0), file=sys.stderr, flush=True)
