"""ctypes loader for the C restatement (oracle/bn254_port.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this.  All arguments and results are ``bytes``/numpy uint8 buffers in gnark memory layout.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libbn254_port.so")
G1_BYTES, G2_BYTES, GT_BYTES, SCALAR_BYTES = 64, 128, 384, 32


def build(force=False):
    src = os.path.join(_HERE, "bn254_port.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _buf(x):
    a = np.frombuffer(x, dtype=np.uint8) if isinstance(x, (bytes, bytearray, memoryview)) else np.ascontiguousarray(x).view(np.uint8).reshape(-1)
    return a, a.ctypes.data_as(ctypes.c_void_p)


def _call2(name, a, b, n, out_bytes, threads):
    out = np.empty(n * out_bytes, dtype=np.uint8)
    a_, ap = _buf(a)
    b_, bp = _buf(b)
    rc = getattr(lib(), name)(ap, bp, ctypes.c_size_t(n), out.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(threads))
    if rc != 0:
        raise ValueError("invalid inputs sizes")
    return out


def _callk(name, a, b, n, k, out_bytes, threads):
    if k == 0:
        raise ValueError("invalid inputs sizes")
    out = np.empty(n * out_bytes, dtype=np.uint8)
    a_, ap = _buf(a)
    b_, bp = _buf(b)
    assert a_.size == n * k * G1_BYTES and b_.size == n * k * G2_BYTES
    rc = getattr(lib(), name)(ap, bp, ctypes.c_size_t(n), ctypes.c_size_t(k), out.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(threads))
    if rc != 0:
        raise ValueError("invalid inputs sizes")
    return out


def pair_batch(P, Q, n, threads=1):
    return _callk("bn254_port_multi_pair_batch", P, Q, n, 1, GT_BYTES, threads)


def multi_pair_batch(P, Q, n, k, threads=1):
    return _callk("bn254_port_multi_pair_batch", P, Q, n, k, GT_BYTES, threads)


def miller_loop_batch(P, Q, n, k, threads=1):
    return _callk("bn254_port_miller_loop_batch", P, Q, n, k, GT_BYTES, threads)


def pairing_check_batch(P, Q, n, k, threads=1):
    return _callk("bn254_port_pairing_check_batch", P, Q, n, k, 1, threads)


def final_exp_batch(f, n, threads=1):
    out = np.empty(n * GT_BYTES, dtype=np.uint8)
    _, fp_ = _buf(f)
    lib().bn254_port_final_exp_batch(fp_, ctypes.c_size_t(n), out.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(threads))
    return out


def _unary(name, f, n, threads=1):
    out = np.empty(n * GT_BYTES, dtype=np.uint8)
    _, fp_ = _buf(f)
    getattr(lib(), name)(fp_, ctypes.c_size_t(n), out.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(threads))
    return out


def gt_sqr_batch(f, n, threads=1):
    return _unary("bn254_port_gt_sqr_batch", f, n, threads)


def gt_cyclo_sqr_batch(f, n, threads=1):
    return _unary("bn254_port_gt_cyclo_sqr_batch", f, n, threads)


def g1_mul_batch(base, s, n, threads=1):
    return _call2("bn254_port_g1_mul_batch", base, s, n, G1_BYTES, threads)


def g2_mul_batch(base, s, n, threads=1):
    return _call2("bn254_port_g2_mul_batch", base, s, n, G2_BYTES, threads)


def g1_mul_base_batch(base, s, n, threads=1):
    return _call2("bn254_port_g1_mul_base_batch", base, s, n, G1_BYTES, threads)


def g2_mul_base_batch(base, s, n, threads=1):
    return _call2("bn254_port_g2_mul_base_batch", base, s, n, G2_BYTES, threads)


def g1_add_batch(a, b, n, threads=1):
    return _call2("bn254_port_g1_add_batch", a, b, n, G1_BYTES, threads)


def g2_add_batch(a, b, n, threads=1):
    return _call2("bn254_port_g2_add_batch", a, b, n, G2_BYTES, threads)


def gt_exp_batch(x, k, n, threads=1):
    return _call2("bn254_port_gt_exp_batch", x, k, n, GT_BYTES, threads)


def gt_exp_base_batch(x, k, n, threads=1):
    return _call2("bn254_port_gt_exp_base_batch", x, k, n, GT_BYTES, threads)


def gt_mul_batch(a, b, n, threads=1):
    return _call2("bn254_port_gt_mul_batch", a, b, n, GT_BYTES, threads)


def gt_div_batch(a, b, n, threads=1):
    return _call2("bn254_port_gt_div_batch", a, b, n, GT_BYTES, threads)


def fp_mul_batch(a, b, n, threads=1):
    return _call2("bn254_port_fp_mul_batch", a, b, n, 32, threads)


def generators():
    g1 = np.empty(G1_BYTES, dtype=np.uint8)
    g2 = np.empty(G2_BYTES, dtype=np.uint8)
    lib().bn254_port_generators(g1.ctypes.data_as(ctypes.c_void_p), g2.ctypes.data_as(ctypes.c_void_p))
    return g1, g2
