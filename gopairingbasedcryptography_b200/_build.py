"""Build libbn254_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

One translation unit per kernel family (csrc/k_*.cu) plus the host runtime (csrc/engine.cu), compiled in parallel and
linked into one shared library.  Each family has its own compile flags: the register cap that is right for the
pairing tower is not the one the group ladders want."""
from __future__ import annotations

import glob
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# Default flags = the measured best of the profiles/r1 sweeps: out-of-line add-type leaves and Montgomery product
# (small instruction footprint), 3 CTAs/SM (168 registers), per-thread shared-memory scratch for the staged tower,
# CTA lockstep barriers (the four warps of a CTA share instruction-cache lines).
DEFAULT = ["-DBN254_OOL_ADDS", "-DBN254_OOL_FPMUL", "-DBN254_MIN_BLOCKS=3", "-DBN254_SMEM_SCRATCH", "-DBN254_CTA_LOCKSTEP"]
NOLS = [f for f in DEFAULT if f != "-DBN254_CTA_LOCKSTEP"]
UNITS = ["k_pairing", "k_group_mul", "k_group_fixed", "k_group_add", "k_gt", "k_hash", "k_vm", "k_wvm", "k_fr", "engine"]
VARIANTS = {  # experiment builds selected with BN254_VARIANT=<name>: flags for every unit, or {unit: flags}
    "": DEFAULT,
    "nolockstep": NOLS,
    "karatsuba_mulx": DEFAULT + ["-DBN254_KARATSUBA_MULX"],
    "ool_jac": DEFAULT + ["-DBN254_OOL_JAC"],
    "wvmprof": DEFAULT + ["-DBN254_WVM_PROFILE"],
    "wvmprof_noinline": DEFAULT + ["-DBN254_WVM_PROFILE"],  # built after `WVM_TINLINE=0 python csrc/wvmgen.py` (program-shape experiment)
    "nosmem": ["-DBN254_OOL_ADDS", "-DBN254_OOL_FPMUL", "-DBN254_MIN_BLOCKS=3"],
    # timing-only probe of 4 CTAs per SM (128 registers, 16 warps): the 9 scratch slots are squeezed into a 400-byte stride, so
    # slots 6-8 alias the next thread's -- results are WRONG, the instruction stream and its timing are those of a design
    # whose three extra slots are as fast as shared memory (upper bound for a TMEM-backed variant)
    "b4_timing_only": {"k_pairing": ["-DBN254_OOL_ADDS", "-DBN254_OOL_FPMUL", "-DBN254_MIN_BLOCKS=4", "-DBN254_SMEM_SCRATCH", "-DBN254_CTA_LOCKSTEP",
                                     "-DBN254_SCRATCH_STRIDE=400"]},
    "b4_timing_only_seq": {"k_pairing": ["-DBN254_OOL_ADDS", "-DBN254_OOL_FPMUL", "-DBN254_MIN_BLOCKS=4", "-DBN254_SMEM_SCRATCH", "-DBN254_CTA_LOCKSTEP",
                                         "-DBN254_SCRATCH_STRIDE=400", "-DBN254_MULX_SEQ", "-DBN254_CYC_LATE_LOADS"]},
    "mulx_seq": {"k_pairing": DEFAULT + ["-DBN254_MULX_SEQ", "-DBN254_CYC_LATE_LOADS"]},
    "wvm_ahead2": {"k_wvm": DEFAULT + ["-DWVM_AHEAD=2"]},
    "wvm_ahead3": {"k_wvm": DEFAULT + ["-DWVM_AHEAD=3"]},
    "wvm_ahead4": {"k_wvm": DEFAULT + ["-DWVM_AHEAD=4"]},
    "wvm_ahead1": {"k_wvm": DEFAULT + ["-DWVM_AHEAD=1"]},
    "wvm_onepass": {"k_wvm": DEFAULT + ["-DWVM_FINISH_ONEPASS=1"]},
    "wvm_linbatch": {"k_wvm": DEFAULT + ["-DWVM_LIN_BATCH=1"]},
    # program-shape experiments: built after `WVM_TINLINE_FINALEXP=<n> python csrc/wvmgen.py` with BN254_KEEP_GENERATED=1
    "fe_tinline12": DEFAULT, "fe_tinline15": DEFAULT,
    "lines16": DEFAULT + ["-DBN254_LINES_CHUNK=16"], "lines12": DEFAULT + ["-DBN254_LINES_CHUNK=12"],
    "pair_add_lines": {"k_pairing": DEFAULT + ["-DBN254_PAIR_ADD_LINES"]},
    "gt_glv2": {"k_gt": DEFAULT + ["-DBN254_GT_EXP_GLV2"]},  # GT-proper exponentiation by the two-dimensional split (A/B of the GLS-4 ladder)
    "euler_ladder": {"k_hash": DEFAULT + ["-DBN254_EULER_LADDER"]},  # quadratic character by the ladder a^((p-1)/2) (A/B of the Jacobi symbol)
    "lines_global": {"k_pairing": DEFAULT + ["-DBN254_LINES_TMA=0"]},  # line tables read straight from global memory (A/B of the TMA ring)
    "b2": ["-DBN254_OOL_ADDS", "-DBN254_OOL_FPMUL", "-DBN254_MIN_BLOCKS=2", "-DBN254_SMEM_SCRATCH", "-DBN254_CTA_LOCKSTEP"],
}
# per-unit additions on top of the variant's flags
UNIT_FLAGS = {}
VARIANT = os.environ.get("BN254_VARIANT", "")
LIB = os.path.join(HERE, "lib", "libbn254_b200%s.so" % ("_" + VARIANT if VARIANT else ""))
OBJDIR = os.path.join(HERE, "lib", "obj" + ("_" + VARIANT if VARIANT else ""))
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
              "-Xcompiler", "-fvisibility=default"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) +
                  glob.glob(os.path.join(CSRC, "*.inc")) + [os.path.join(HERE, "..", "include", "bn254_b200.h"), os.path.abspath(__file__)])


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in sources())


def unit_flags(unit):
    v = VARIANTS[VARIANT]
    flags = list(v.get(unit, v.get("*", DEFAULT)) if isinstance(v, dict) else v)
    return flags + UNIT_FLAGS.get(unit, []) + os.environ.get("BN254_EXTRA_FLAGS", "").split()


def ensure_generated():
    if os.environ.get("BN254_KEEP_GENERATED"):
        return
    """The warp-VM programs (wvm_prog_*.inc, ~8 MB of text) are generated, not committed: wvmgen.py rebuilds them from
    the traced pairing formulas in a few seconds."""
    outs = [os.path.join(CSRC, f) for f in ("wvm_prog_miller.inc", "wvm_prog_miller2.inc", "wvm_prog_finalexp.inc", "wvm_prog_meta.cuh")]
    srcs = [os.path.join(CSRC, f) for f in ("wvmgen.py", "vmgen.py")]
    if all(os.path.exists(o) for o in outs) and min(os.path.getmtime(o) for o in outs) >= max(os.path.getmtime(s) for s in srcs):
        return
    import sys

    subprocess.check_call([sys.executable, os.path.join(CSRC, "wvmgen.py")])


def build(force=False, verbose=False):
    ensure_generated()
    if not force and not is_stale():
        return LIB
    os.makedirs(OBJDIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

    def compile_unit(unit):
        obj = os.path.join(OBJDIR, unit + ".o")
        cmd = [nvcc] + NVCC_FLAGS + unit_flags(unit) + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, os.path.join(CSRC, unit + ".cu")]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (unit, r.stdout, r.stderr))
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(len(UNITS), os.cpu_count() or 4)) as ex:
        objs = list(ex.map(compile_unit, UNITS))
    subprocess.check_call([nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"])
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
