"""ctypes binding of include/bn254_b200.h.  Fails loudly if the CUDA library is missing."""
from __future__ import annotations

import ctypes
import os

from ._build import LIB

_lib = None

# every function include/bn254_b200.h declares (tests/test_abi_and_host.py checks the list against the header)
SYMBOLS = [
    "bn254_ctx_create", "bn254_ctx_destroy", "bn254_last_error", "bn254_device_count", "bn254_host_alloc",
    "bn254_host_free", "bn254_launch_count", "bn254_sm_count", "bn254_dev_alloc", "bn254_dev_free",
    "bn254_dev_upload", "bn254_dev_download", "bn254_stream_create", "bn254_stream_destroy", "bn254_stream_sync",
    "bn254_generators", "bn254_pair_batch", "bn254_pair_batch_dev", "bn254_multi_pair_batch",
    "bn254_multi_pair_batch_dev", "bn254_g2_lines_create", "bn254_g2_lines_destroy", "bn254_g2_lines_count",
    "bn254_multi_pair_lines_batch", "bn254_multi_pair_lines_batch_dev", "bn254_pairing_check_batch",
    "bn254_pairing_check_batch_dev", "bn254_miller_loop_batch", "bn254_final_exp_batch",
    "bn254_miller_loop_batch_dev", "bn254_final_exp_batch_dev", "bn254_g1_mul_batch", "bn254_g2_mul_batch",
    "bn254_g1_mul_base_batch", "bn254_g2_mul_base_batch", "bn254_g1_mul_batch_dev", "bn254_g2_mul_batch_dev",
    "bn254_fixed_base_create", "bn254_fixed_base_destroy", "bn254_fixed_base_group", "bn254_g1_fixed_mul_batch",
    "bn254_g2_fixed_mul_batch", "bn254_gt_fixed_exp_batch", "bn254_g1_fixed_mul_batch_dev",
    "bn254_g2_fixed_mul_batch_dev", "bn254_gt_fixed_exp_batch_dev", "bn254_msm_table_create",
    "bn254_msm_table_destroy", "bn254_msm_table_len", "bn254_msm_batch", "bn254_msm_batch_dev", "bn254_g1_add_batch",
    "bn254_g2_add_batch", "bn254_g1_add_batch_dev", "bn254_g2_add_batch_dev", "bn254_g1_neg_batch_dev",
    "bn254_g2_neg_batch_dev", "bn254_g1_subset_sum_batch", "bn254_g2_subset_sum_batch",
    "bn254_g1_subset_sum_batch_dev", "bn254_g2_subset_sum_batch_dev", "bn254_g1_sum_batch", "bn254_g2_sum_batch",
    "bn254_g1_sum_batch_dev", "bn254_g2_sum_batch_dev", "bn254_gt_exp_batch", "bn254_gt_exp_base_batch",
    "bn254_gt_exp_batch_dev", "bn254_gt_cyclo_exp_batch", "bn254_gt_cyclo_exp_base_batch",
    "bn254_gt_cyclo_exp_batch_dev", "bn254_gt_mul_batch", "bn254_gt_div_batch", "bn254_gt_mul_batch_dev",
    "bn254_gt_div_batch_dev", "bn254_gt_cyclo_div_batch", "bn254_gt_cyclo_div_batch_dev", "bn254_pairing_check2_fixed_g1_batch", "bn254_pairing_check2_fixed_g1_batch_dev",
    "bn254_hash_to_g1_batch", "bn254_hash_to_g2_batch", "bn254_hash_to_g1_batch_dev", "bn254_hash_to_g2_batch_dev",
    "bn254_fr_lagrange_basis", "bn254_fr_poly_from_roots", "bn254_fr_poly_from_roots_dev",
    "bn254_fr_quotient_coeffs", "bn254_fr_quotient_coeffs_dev", "bn254_fr_to_scalars_dev", "bn254_fr_to_scalars",
    "bn254_fp_mul_batch",
]


def lib():
    """Load libbn254_b200.so (built by __graft_entry__.build()).  No fallback of any kind."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            raise RuntimeError(
                "libbn254_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'`; "
                "this package has no CPU path." % LIB)
        L = ctypes.CDLL(LIB)
        L.bn254_ctx_create.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]
        L.bn254_ctx_create.restype = ctypes.c_int
        L.bn254_ctx_destroy.argtypes = [ctypes.c_void_p]
        L.bn254_ctx_destroy.restype = None
        L.bn254_last_error.argtypes = [ctypes.c_void_p]
        L.bn254_last_error.restype = ctypes.c_char_p
        L.bn254_launch_count.argtypes = [ctypes.c_void_p]
        L.bn254_launch_count.restype = ctypes.c_uint64
        L.bn254_host_alloc.argtypes = [ctypes.c_size_t]
        L.bn254_host_alloc.restype = ctypes.c_void_p
        L.bn254_host_free.argtypes = [ctypes.c_void_p]
        L.bn254_host_free.restype = None
        L.bn254_generators.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.bn254_generators.restype = None
        _lib = L
    return _lib
