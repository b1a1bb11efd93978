"""Build libbn254_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import glob
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# Default flags = the measured best of the profiles/r1 sweeps: out-of-line add-type leaves and Montgomery product
# (small instruction footprint), 3 CTAs/SM (168 registers), per-thread shared-memory scratch for the staged tower,
# CTA lockstep barriers (the four warps of a CTA share instruction-cache lines).
DEFAULT = ["-DBN254_OOL_ADDS", "-DBN254_OOL_FPMUL", "-DBN254_MIN_BLOCKS=3", "-DBN254_SMEM_SCRATCH", "-DBN254_CTA_LOCKSTEP"]
NOLS = [f for f in DEFAULT if f != "-DBN254_CTA_LOCKSTEP"]
VARIANTS = {  # experiment builds selected with BN254_VARIANT=<name>
    "": DEFAULT,
    "nolockstep": NOLS,
    "blk96": [f for f in DEFAULT if "MIN_BLOCKS" not in f] + ["-DBN254_MIN_BLOCKS=4", "-DBN254_BLOCK=96"],
    "blk192": [f for f in DEFAULT if "MIN_BLOCKS" not in f] + ["-DBN254_MIN_BLOCKS=2", "-DBN254_BLOCK=192"],
    "blk64": [f for f in DEFAULT if "MIN_BLOCKS" not in f] + ["-DBN254_MIN_BLOCKS=6", "-DBN254_BLOCK=64"],
    "karatsuba_mulx": DEFAULT + ["-DBN254_KARATSUBA_MULX"],
    "ool_jac": DEFAULT + ["-DBN254_OOL_JAC"],
    "inline_fpmul": [f for f in DEFAULT if f != "-DBN254_OOL_FPMUL"],
    "nosmem": ["-DBN254_OOL_ADDS", "-DBN254_OOL_FPMUL", "-DBN254_MIN_BLOCKS=3"],
    "b2": ["-DBN254_OOL_ADDS", "-DBN254_OOL_FPMUL", "-DBN254_MIN_BLOCKS=2", "-DBN254_SMEM_SCRATCH", "-DBN254_CTA_LOCKSTEP"],
    **{"k%dw%d" % (k, w): DEFAULT + ["-DBN254_VM_K=%d" % k, "-DBN254_VM_WARPS=%d" % w]
       for k in (1, 2, 3, 4, 6) for w in (1, 2, 3, 4, 6, 8)},
}
VARIANT = os.environ.get("BN254_VARIANT", "")
LIB = os.path.join(HERE, "lib", "libbn254_b200%s.so" % ("_" + VARIANT if VARIANT else ""))
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh")) +
                  glob.glob(os.path.join(CSRC, "*.inc")) + [os.path.join(HERE, "..", "include", "bn254_b200.h")])


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force=False, verbose=False):
    if not force and not is_stale():
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + VARIANTS[VARIANT] + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB, os.path.join(CSRC, "engine.cu")]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
