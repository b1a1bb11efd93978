"""Host-side mirror of the gnark-crypto ``ecc/bn254`` call surface used by the reference schemes,
bound to the CUDA engine through the C ABI (include/bn254_b200.h).

Two levels:

* :class:`Engine` -- batch entry points on raw buffers in gnark memory layout (numpy ``uint8``
  arrays or ``bytes``): ``pair_batch``, ``multi_pair_batch``, ``pairing_check_batch``,
  ``miller_loop_batch``, ``final_exp_batch``, ``g1_mul_batch`` ... These are what the Go package's
  ``PairBatch`` etc. bind.
* gnark-named values and functions -- ``G1Affine``, ``G2Affine``, ``GT``, ``Pair``, ``PairingCheck``,
  ``MillerLoop``, ``FinalExponentiation``, ``Generators`` -- with the reference's argument meaning
  and error behaviour (``ValueError("invalid inputs sizes")`` where gnark returns that error;
  nothing else validates its input).  Each call is a 1-element batch on the GPU: they exist so the
  parity tests read like the reference's tests, not for throughput.

Reference call sites: bn254.Pair access/tree/access_tree_node.go:106,110; PairingCheck
signature/bls01_signature/bls_signature.go:81-84; ScalarMultiplication[Base]
signature/bls01_signature/bls_signature.go:45,63; GT.Exp access/tree/access_tree_node.go:156.
No CPU path exists: without the built CUDA library and a B200 every call raises.
"""
from __future__ import annotations

import ctypes
import threading

import numpy as np

from . import _native

P_MOD = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
G1_BYTES, G2_BYTES, GT_BYTES, SCALAR_BYTES = 64, 128, 384, 32

_ERRORS = {-1: "invalid inputs sizes", -2: "CUDA error", -3: "out of memory", -4: "bad argument"}


class EngineError(RuntimeError):
    pass


def _u8(x, item_bytes, what):
    if isinstance(x, (bytes, bytearray, memoryview)):
        a = np.frombuffer(x, dtype=np.uint8)
    else:
        a = np.ascontiguousarray(x)
        a = a.view(np.uint8).reshape(-1)
    if a.size % item_bytes:
        raise ValueError("%s: buffer of %d bytes is not a multiple of %d" % (what, a.size, item_bytes))
    return a


def scalars_to_bytes(ks):
    """ints -> (n,32) little-endian regular-form scalars (the C-ABI scalar format)."""
    out = np.empty((len(ks), SCALAR_BYTES), dtype=np.uint8)
    for i, k in enumerate(ks):
        k = int(k)
        if k < 0 or k >> 256:
            raise ValueError("scalar out of range [0, 2^256)")
        out[i] = np.frombuffer(k.to_bytes(32, "little"), dtype=np.uint8)
    return out


class Engine:
    """One context per GPU (bn254_ctx).  Thread-safe; buffers are copied, nothing is retained."""

    def __init__(self, device=0):
        self._lib = _native.lib()
        h = ctypes.c_void_p()
        rc = self._lib.bn254_ctx_create(int(device), ctypes.byref(h))
        if rc != 0:
            raise EngineError("bn254_ctx_create(device=%d) failed: %s (this engine has no CPU fallback)"
                              % (device, _ERRORS.get(rc, rc)))
        self._h = h
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.bn254_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    @property
    def launches(self):
        return int(self._lib.bn254_launch_count(self._h))

    def _check(self, rc):
        if rc == 0:
            return
        if rc == -1:
            raise ValueError("invalid inputs sizes")
        raise EngineError("%s: %s" % (_ERRORS.get(rc, rc), self._lib.bn254_last_error(self._h).decode()))

    def _call(self, name, bufs, sizes, out_bytes, n, out_dtype=np.uint8, pre_sizes=None, mid_sizes=None, out=None):
        out = self._out(out, n * out_bytes)
        fn = getattr(self._lib, name)
        if pre_sizes is not None:  # (ctx, buf0, size..., buf1, size..., out) argument order
            args = [self._h, bufs[0].ctypes.data_as(ctypes.c_void_p)] + [ctypes.c_size_t(s) for s in pre_sizes]
            args += [bufs[1].ctypes.data_as(ctypes.c_void_p)] + [ctypes.c_size_t(s) for s in mid_sizes]
        else:
            args = [self._h] + [b.ctypes.data_as(ctypes.c_void_p) for b in bufs] + [ctypes.c_size_t(s) for s in sizes]
        args.append(out.ctypes.data_as(ctypes.c_void_p))
        fn.restype = ctypes.c_int
        self._check(fn(*args))
        return out

    @staticmethod
    def _out(out, nbytes):
        """Result buffer: a fresh array, or the caller's (`out=`; page-locked memory -- bn254_host_alloc,
        torch.Tensor.pin_memory -- is filled by direct device->host copies, no staging memcpy, no page faults)."""
        if out is None:
            return np.empty(nbytes, dtype=np.uint8)
        if not isinstance(out, np.ndarray) or not out.flags["C_CONTIGUOUS"] or not out.flags["WRITEABLE"]:
            raise ValueError("out must be a writable C-contiguous numpy array (a strided view would be copied, not filled)")
        out = out.reshape(-1).view(np.uint8)
        if out.size != nbytes:
            raise ValueError("invalid inputs sizes")
        return out

    # ---- pairings -------------------------------------------------------------------------
    def pair_batch(self, P, Q, out=None):
        """n independent pairings e(P[i], Q[i]) -> (n, 384).  out: optional caller-owned (n, 384) uint8 buffer; with
        P, Q and out in page-locked memory the library copies to / from the device directly (no staging memcpy)."""
        P, Q = _u8(P, G1_BYTES, "P"), _u8(Q, G2_BYTES, "Q")
        n = P.size // G1_BYTES
        if n != Q.size // G2_BYTES:
            raise ValueError("invalid inputs sizes")
        return self._call("bn254_pair_batch", [P, Q], [n], GT_BYTES, n, out=out).reshape(n, GT_BYTES)

    def _kpairs(self, name, P, Q, k, out_bytes):
        P, Q = _u8(P, G1_BYTES, "P"), _u8(Q, G2_BYTES, "Q")
        k = int(k)
        if k <= 0 or P.size // G1_BYTES != Q.size // G2_BYTES or (P.size // G1_BYTES) % k:
            raise ValueError("invalid inputs sizes")
        n = P.size // G1_BYTES // k
        return self._call(name, [P, Q], [n, k], out_bytes, n), n

    def multi_pair_batch(self, P, Q, k):
        """n products of k pairings with ONE final exponentiation each -> (n, 384)."""
        out, n = self._kpairs("bn254_multi_pair_batch", P, Q, k, GT_BYTES)
        return out.reshape(n, GT_BYTES)

    def miller_loop_batch(self, P, Q, k=1):
        out, n = self._kpairs("bn254_miller_loop_batch", P, Q, k, GT_BYTES)
        return out.reshape(n, GT_BYTES)

    def pairing_check_batch(self, P, Q, k):
        out, n = self._kpairs("bn254_pairing_check_batch", P, Q, k, 1)
        return out.astype(bool)

    def g2_lines_create(self, Q):
        """Precompute the Miller-loop line tables of m fixed G2 points (kept on the GPU); returns a handle."""
        Q = _u8(Q, G2_BYTES, "Q")
        m = Q.size // G2_BYTES
        h = ctypes.c_void_p()
        fn = self._lib.bn254_g2_lines_create
        fn.restype = ctypes.c_int
        self._check(fn(self._h, Q.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(m), ctypes.byref(h)))
        return G2Lines(self, h, m)

    def multi_pair_lines_batch(self, P, lines):
        """n products of m pairings against the m table points: P is (n*m, 64) -> (n, 384)."""
        P = _u8(P, G1_BYTES, "P")
        m = lines.m
        if lines.engine is not self or (P.size // G1_BYTES) % m:
            raise ValueError("invalid inputs sizes")
        n = P.size // G1_BYTES // m
        out = np.empty(n * GT_BYTES, dtype=np.uint8)
        fn = self._lib.bn254_multi_pair_lines_batch
        fn.restype = ctypes.c_int
        self._check(fn(self._h, P.ctypes.data_as(ctypes.c_void_p), lines.handle, ctypes.c_size_t(n), out.ctypes.data_as(ctypes.c_void_p)))
        return out.reshape(n, GT_BYTES)

    def final_exp_batch(self, f):
        f = _u8(f, GT_BYTES, "f")
        n = f.size // GT_BYTES
        return self._call("bn254_final_exp_batch", [f], [n], GT_BYTES, n).reshape(n, GT_BYTES)

    # ---- groups ---------------------------------------------------------------------------
    def _mul(self, name, base, scalars, pt_bytes, broadcast, out=None):
        base, scalars = _u8(base, pt_bytes, "base"), _u8(scalars, SCALAR_BYTES, "scalars")
        n = scalars.size // SCALAR_BYTES
        if (broadcast and base.size != pt_bytes) or (not broadcast and base.size // pt_bytes != n):
            raise ValueError("invalid inputs sizes")
        return self._call(name, [base, scalars], [n], pt_bytes, n, out=out).reshape(n, pt_bytes)

    def g1_mul_batch(self, base, scalars, out=None):
        return self._mul("bn254_g1_mul_batch", base, scalars, G1_BYTES, False, out)

    def g2_mul_batch(self, base, scalars, out=None):
        return self._mul("bn254_g2_mul_batch", base, scalars, G2_BYTES, False, out)

    def g1_mul_base_batch(self, base1, scalars, out=None):
        return self._mul("bn254_g1_mul_base_batch", base1, scalars, G1_BYTES, True, out)

    def g2_mul_base_batch(self, base1, scalars, out=None):
        return self._mul("bn254_g2_mul_base_batch", base1, scalars, G2_BYTES, True, out)

    def _binary(self, name, a, b, item, out=None):
        a, b = _u8(a, item, "a"), _u8(b, item, "b")
        if a.size != b.size:
            raise ValueError("invalid inputs sizes")
        n = a.size // item
        return self._call(name, [a, b], [n], item, n, out=out).reshape(n, item)

    def g1_add_batch(self, a, b):
        return self._binary("bn254_g1_add_batch", a, b, G1_BYTES)

    def g2_add_batch(self, a, b):
        return self._binary("bn254_g2_add_batch", a, b, G2_BYTES)

    def _subset_sum(self, name, U, sel, pt_bytes):
        U = _u8(U, pt_bytes, "U")
        m = U.size // pt_bytes - 1
        if m <= 0:
            raise ValueError("invalid inputs sizes")
        row = (m + 7) // 8
        sel = _u8(sel, row, "sel")
        n = sel.size // row
        return self._call(name, [U, sel], [], pt_bytes, n, pre_sizes=[m], mid_sizes=[n]).reshape(n, pt_bytes)

    def g1_subset_sum_batch(self, U, sel):
        """out[i] = U[0] + sum_{bit j of sel[i]} U[j+1]; sel rows are MSB-first bit strings of ceil(m/8) bytes."""
        return self._subset_sum("bn254_g1_subset_sum_batch", U, sel, G1_BYTES)

    def g2_subset_sum_batch(self, U, sel):
        return self._subset_sum("bn254_g2_subset_sum_batch", U, sel, G2_BYTES)

    def _segment_sum(self, name, pts, length, pt_bytes):
        pts = _u8(pts, pt_bytes, "points")
        length = int(length)
        if length <= 0 or (pts.size // pt_bytes) % length:
            raise ValueError("invalid inputs sizes")
        groups = pts.size // pt_bytes // length
        return self._call(name, [pts], [groups, length], pt_bytes, groups).reshape(groups, pt_bytes)

    def g1_sum_batch(self, pts, length):
        """Sum every `length` consecutive points -> (groups, 64)."""
        return self._segment_sum("bn254_g1_sum_batch", pts, length, G1_BYTES)

    def g2_sum_batch(self, pts, length):
        return self._segment_sum("bn254_g2_sum_batch", pts, length, G2_BYTES)

    # ---- GT -------------------------------------------------------------------------------
    def gt_exp_batch(self, x, k, out=None):
        x, k = _u8(x, GT_BYTES, "x"), _u8(k, SCALAR_BYTES, "k")
        n = k.size // SCALAR_BYTES
        if x.size // GT_BYTES != n:
            raise ValueError("invalid inputs sizes")
        return self._call("bn254_gt_exp_batch", [x, k], [n], GT_BYTES, n, out=out).reshape(n, GT_BYTES)

    def gt_exp_base_batch(self, x1, k, out=None):
        x1, k = _u8(x1, GT_BYTES, "x"), _u8(k, SCALAR_BYTES, "k")
        n = k.size // SCALAR_BYTES
        if x1.size != GT_BYTES:
            raise ValueError("invalid inputs sizes")
        return self._call("bn254_gt_exp_base_batch", [x1, k], [n], GT_BYTES, n, out=out).reshape(n, GT_BYTES)

    def gt_cyclo_exp_batch(self, x, k, out=None):
        """x[i]^k[i] for x in the cyclotomic subgroup (pairing outputs and their products/powers)."""
        x, k = _u8(x, GT_BYTES, "x"), _u8(k, SCALAR_BYTES, "k")
        n = k.size // SCALAR_BYTES
        if x.size // GT_BYTES != n:
            raise ValueError("invalid inputs sizes")
        return self._call("bn254_gt_cyclo_exp_batch", [x, k], [n], GT_BYTES, n, out=out).reshape(n, GT_BYTES)

    def gt_cyclo_exp_base_batch(self, x1, k, out=None):
        x1, k = _u8(x1, GT_BYTES, "x"), _u8(k, SCALAR_BYTES, "k")
        n = k.size // SCALAR_BYTES
        if x1.size != GT_BYTES:
            raise ValueError("invalid inputs sizes")
        return self._call("bn254_gt_cyclo_exp_base_batch", [x1, k], [n], GT_BYTES, n, out=out).reshape(n, GT_BYTES)

    def gt_mul_batch(self, a, b, out=None):
        return self._binary("bn254_gt_mul_batch", a, b, GT_BYTES, out)

    def gt_div_batch(self, a, b, out=None):
        return self._binary("bn254_gt_div_batch", a, b, GT_BYTES, out)

    def gt_cyclo_div_batch(self, a, b, out=None):
        """a[i] / b[i] for b in GT proper (pairing outputs and their products / powers): b^-1 = conj(b), one product."""
        return self._binary("bn254_gt_cyclo_div_batch", a, b, GT_BYTES, out)

    def fp_mul_batch(self, a, b):
        return self._binary("bn254_fp_mul_batch", a, b, 32)

    def pairing_check2_fixed_g1_batch(self, p0, p1, q0, q1, out=None):
        """ok[i] = PairingCheck([p0, p1], [q0[i], q1[i]]): two G1 points shared by the batch (BLS verification shape).
        out: optional caller-owned (n,) uint8 buffer for the 0/1 results (returned as a bool view)."""
        p01 = np.concatenate([_u8(p0, G1_BYTES, "p0"), _u8(p1, G1_BYTES, "p1")])
        q0, q1 = _u8(q0, G2_BYTES, "q0"), _u8(q1, G2_BYTES, "q1")
        if p01.size != 2 * G1_BYTES or q0.size != q1.size:
            raise ValueError("invalid inputs sizes")
        n = q0.size // G2_BYTES
        out = self._out(out, n)
        fn = self._lib.bn254_pairing_check2_fixed_g1_batch
        fn.restype = ctypes.c_int
        self._check(fn(self._h, p01.ctypes.data_as(ctypes.c_void_p), q0.ctypes.data_as(ctypes.c_void_p),
                       q1.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(n), out.ctypes.data_as(ctypes.c_void_p)))
        return out.view(bool)

    # ---- hash-to-curve (gnark bn254.HashToG1 / HashToG2; hash/hash_to.go in the reference) ----
    def _hash_to_curve(self, name, msgs, dst, out_bytes):
        """msgs: a sequence of bytes-like messages, or a pair (blob, offsets) — the concatenated messages as uint8 and
        n + 1 uint64 offsets — which is what the C ABI takes and skips the per-message Python work."""
        dst = bytes(dst)
        if isinstance(msgs, tuple) and len(msgs) == 2 and isinstance(msgs[1], np.ndarray):
            blob = np.ascontiguousarray(msgs[0], dtype=np.uint8).reshape(-1)
            off = np.ascontiguousarray(msgs[1], dtype=np.uint64).reshape(-1)
            n = off.size - 1
            if n < 0 or (np.diff(off.astype(np.int64)) < 0).any() or (n >= 0 and int(off[-1]) > blob.size):
                raise ValueError("invalid message offsets")
            if blob.size == 0:
                blob = np.zeros(1, dtype=np.uint8)
        else:
            if not isinstance(msgs, (list, tuple)):
                msgs = list(msgs)
            n = len(msgs)
            off = np.zeros(n + 1, dtype=np.uint64)
            if n:
                np.cumsum(np.fromiter(map(len, msgs), dtype=np.uint64, count=n), out=off[1:])
            blob = np.frombuffer(b"".join(msgs) or b"\0", dtype=np.uint8)
        d = np.frombuffer(dst or b"\0", dtype=np.uint8)
        out = np.empty(n * out_bytes, dtype=np.uint8)
        fn = getattr(self._lib, name)
        fn.restype = ctypes.c_int
        self._check(fn(self._h, blob.ctypes.data_as(ctypes.c_void_p), off.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(n),
                       d.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(len(dst)), out.ctypes.data_as(ctypes.c_void_p)))
        return out.reshape(n, out_bytes)

    def hash_to_g1_batch(self, msgs, dst):
        """n x bn254.HashToG1(msg, dst) -> (n, 64) affine points."""
        return self._hash_to_curve("bn254_hash_to_g1_batch", msgs, dst, G1_BYTES)

    def hash_to_g2_batch(self, msgs, dst):
        """n x bn254.HashToG2(msg, dst) -> (n, 128) affine points."""
        return self._hash_to_curve("bn254_hash_to_g2_batch", msgs, dst, G2_BYTES)


    # ---- explicit table handles ------------------------------------------------------------
    def fixed_base_create(self, group, base):
        """Immutable 32 x 255 window table of ONE base: group 1 = G1, 2 = G2, 3 = GT (bn254_fixed_base_create)."""
        item = {1: G1_BYTES, 2: G2_BYTES, 3: GT_BYTES}[int(group)]
        base = _u8(base, item, "base")
        if base.size != item:
            raise ValueError("invalid inputs sizes")
        h = ctypes.c_void_p()
        fn = self._lib.bn254_fixed_base_create
        fn.restype = ctypes.c_int
        self._check(fn(self._h, ctypes.c_int(int(group)), base.ctypes.data_as(ctypes.c_void_p), ctypes.byref(h)))
        return FixedBase(self, h, int(group))

    def _fixed(self, name, table, scalars, item, group, out=None):
        if table.engine is not self or table.group != group:
            raise ValueError("fixed-base handle of another engine or group")
        scalars = _u8(scalars, SCALAR_BYTES, "scalars")
        n = scalars.size // SCALAR_BYTES
        out = self._out(out, n * item)
        fn = getattr(self._lib, name)
        fn.restype = ctypes.c_int
        self._check(fn(self._h, table.handle, scalars.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(n), out.ctypes.data_as(ctypes.c_void_p)))
        return out.reshape(n, item)

    def g1_fixed_mul_batch(self, table, scalars, out=None):
        return self._fixed("bn254_g1_fixed_mul_batch", table, scalars, G1_BYTES, 1, out)

    def g2_fixed_mul_batch(self, table, scalars, out=None):
        return self._fixed("bn254_g2_fixed_mul_batch", table, scalars, G2_BYTES, 2, out)

    def gt_fixed_exp_batch(self, table, k, out=None):
        return self._fixed("bn254_gt_fixed_exp_batch", table, k, GT_BYTES, 3, out)

    def msm_table_create(self, group, points):
        """Per-point byte-window tables of `len` shared points (bn254_msm_table_create): group 1 = G1, 2 = G2."""
        item = {1: G1_BYTES, 2: G2_BYTES}[int(group)]
        points = _u8(points, item, "points")
        n = points.size // item
        h = ctypes.c_void_p()
        fn = self._lib.bn254_msm_table_create
        fn.restype = ctypes.c_int
        self._check(fn(self._h, ctypes.c_int(int(group)), points.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(n), ctypes.byref(h)))
        return MsmTable(self, h, int(group), n)

    def msm_batch(self, table, scalars):
        """out[v] = sum_j [scalars[v, j]] P_j over the table's points: scalars (nvec, len, 32) -> (nvec, 64 | 128)."""
        if table.engine is not self:
            raise ValueError("MSM table of another engine")
        scalars = _u8(scalars, SCALAR_BYTES * table.len, "scalars")
        nvec = scalars.size // (SCALAR_BYTES * table.len)
        item = G1_BYTES if table.group == 1 else G2_BYTES
        out = np.empty(nvec * item, dtype=np.uint8)
        fn = self._lib.bn254_msm_batch
        fn.restype = ctypes.c_int
        self._check(fn(self._h, table.handle, scalars.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(nvec), out.ctypes.data_as(ctypes.c_void_p)))
        return out.reshape(nvec, item)

    # ---- Fr feeders (fr.Element = 32 B Montgomery, gnark layout) ------------------------------
    def fr_poly_from_roots(self, roots):
        """Coefficients c_0..c_n (fr.Element) of prod (X - root_i) (afp25_bibe_utils.go:14-43 computePolynomialCoeffs)."""
        roots = _u8(roots, 32, "roots")
        n = roots.size // 32
        out = np.empty((n + 1) * 32, dtype=np.uint8)
        fn = self._lib.bn254_fr_poly_from_roots
        fn.restype = ctypes.c_int
        self._check(fn(self._h, roots.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(n), out.ctypes.data_as(ctypes.c_void_p)))
        return out.reshape(n + 1, 32)

    def fr_quotient_coeffs(self, f, ids):
        """For every id the n coefficients of f(X) / (X - id), as regular-form scalars (nvec, n, 32)."""
        f, ids = _u8(f, 32, "f"), _u8(ids, 32, "ids")
        n, nvec = f.size // 32 - 1, ids.size // 32
        if n <= 0:
            raise ValueError("invalid inputs sizes")
        out = np.empty(nvec * n * 32, dtype=np.uint8)
        fn = self._lib.bn254_fr_quotient_coeffs
        fn.restype = ctypes.c_int
        self._check(fn(self._h, f.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(n), ids.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(nvec),
                       out.ctypes.data_as(ctypes.c_void_p)))
        return out.reshape(nvec, n, 32)

    # ---- device-resident entry points (pointers are ints, e.g. torch.Tensor.data_ptr()) ----
    # kinds: p = device pointer, z = size_t, h = table handle object, b = host bytes (pointer + length)
    _DEV_SIGS = {
        "pair_batch_dev": "ppzp", "multi_pair_batch_dev": "ppzzp", "pairing_check_batch_dev": "ppzzp", "miller_loop_batch_dev": "ppzzp",
        "final_exp_batch_dev": "pzp", "multi_pair_lines_batch_dev": "phzp", "g1_mul_batch_dev": "pzpzp", "g2_mul_batch_dev": "pzpzp",
        "g1_fixed_mul_batch_dev": "hpzp", "g2_fixed_mul_batch_dev": "hpzp", "gt_fixed_exp_batch_dev": "hpzp", "msm_batch_dev": "hpzp",
        "g1_add_batch_dev": "ppzp", "g2_add_batch_dev": "ppzp", "g1_neg_batch_dev": "pzp", "g2_neg_batch_dev": "pzp",
        "g1_subset_sum_batch_dev": "pzpzp", "g2_subset_sum_batch_dev": "pzpzp", "g1_sum_batch_dev": "pzzp", "g2_sum_batch_dev": "pzzp",
        "gt_exp_batch_dev": "pzpzp", "gt_cyclo_exp_batch_dev": "pzpzp", "gt_mul_batch_dev": "pzpzzp", "gt_div_batch_dev": "pzpzzp", "gt_cyclo_div_batch_dev": "pzpzzp",
        "pairing_check2_fixed_g1_batch_dev": "pppzp", "hash_to_g1_batch_dev": "ppzbp", "hash_to_g2_batch_dev": "ppzbp",
        "fr_poly_from_roots_dev": "pzp", "fr_quotient_coeffs_dev": "pzpzp", "fr_to_scalars_dev": "pzp",
    }

    def dev(self, name, *args, stream=0):
        """Enqueue bn254_<name> on `stream` (a cudaStream_t as int); returns immediately, nothing is synchronised."""
        sig = self._DEV_SIGS[name]
        if len(args) != len(sig):
            raise TypeError("%s takes %d arguments" % (name, len(sig)))
        cargs, keep = [self._h], []
        for kind, a in zip(sig, args):
            if kind == "p":
                cargs.append(ctypes.c_void_p(int(a)))
            elif kind == "z":
                cargs.append(ctypes.c_size_t(int(a)))
            elif kind == "h":
                if a.engine is not self:
                    raise ValueError("table handle of another engine")
                cargs.append(a.handle)
            else:
                b = bytes(a)
                buf = ctypes.create_string_buffer(b or b"\0", max(len(b), 1))
                keep.append(buf)
                cargs += [ctypes.cast(buf, ctypes.c_void_p), ctypes.c_size_t(len(b))]
        cargs.append(ctypes.c_void_p(int(stream)))
        fn = getattr(self._lib, "bn254_" + name)
        fn.restype = ctypes.c_int
        self._check(fn(*cargs))

    def _dev(self, name, *args):
        fn = getattr(self._lib, name)
        fn.restype = ctypes.c_int
        self._check(fn(self._h, *args))

    def pair_batch_dev(self, dP, dQ, n, d_out, stream=0):
        self._dev("bn254_pair_batch_dev", ctypes.c_void_p(dP), ctypes.c_void_p(dQ), ctypes.c_size_t(n),
                  ctypes.c_void_p(d_out), ctypes.c_void_p(stream))

    def multi_pair_batch_dev(self, dP, dQ, n, k, d_out, stream=0):
        self._dev("bn254_multi_pair_batch_dev", ctypes.c_void_p(dP), ctypes.c_void_p(dQ), ctypes.c_size_t(n),
                  ctypes.c_size_t(k), ctypes.c_void_p(d_out), ctypes.c_void_p(stream))

    def pairing_check_batch_dev(self, dP, dQ, n, k, d_ok, stream=0):
        self._dev("bn254_pairing_check_batch_dev", ctypes.c_void_p(dP), ctypes.c_void_p(dQ), ctypes.c_size_t(n),
                  ctypes.c_size_t(k), ctypes.c_void_p(d_ok), ctypes.c_void_p(stream))

    def miller_loop_batch_dev(self, dP, dQ, n, k, d_out, stream=0):
        self._dev("bn254_miller_loop_batch_dev", ctypes.c_void_p(dP), ctypes.c_void_p(dQ), ctypes.c_size_t(n),
                  ctypes.c_size_t(k), ctypes.c_void_p(d_out), ctypes.c_void_p(stream))

    def final_exp_batch_dev(self, d_in, n, d_out, stream=0):
        self._dev("bn254_final_exp_batch_dev", ctypes.c_void_p(d_in), ctypes.c_size_t(n), ctypes.c_void_p(d_out),
                  ctypes.c_void_p(stream))

    def g1_mul_batch_dev(self, d_base, stride, d_s, n, d_out, stream=0):
        self._dev("bn254_g1_mul_batch_dev", ctypes.c_void_p(d_base), ctypes.c_size_t(stride), ctypes.c_void_p(d_s),
                  ctypes.c_size_t(n), ctypes.c_void_p(d_out), ctypes.c_void_p(stream))

    def g2_mul_batch_dev(self, d_base, stride, d_s, n, d_out, stream=0):
        self._dev("bn254_g2_mul_batch_dev", ctypes.c_void_p(d_base), ctypes.c_size_t(stride), ctypes.c_void_p(d_s),
                  ctypes.c_size_t(n), ctypes.c_void_p(d_out), ctypes.c_void_p(stream))

    def gt_exp_batch_dev(self, d_x, stride, d_k, n, d_out, stream=0):
        self._dev("bn254_gt_exp_batch_dev", ctypes.c_void_p(d_x), ctypes.c_size_t(stride), ctypes.c_void_p(d_k),
                  ctypes.c_size_t(n), ctypes.c_void_p(d_out), ctypes.c_void_p(stream))


class _Handle:
    _destroy = None

    def close(self):
        if self.handle:
            fn = getattr(self.engine._lib, self._destroy)
            fn.argtypes = [ctypes.c_void_p]
            fn.restype = None
            fn(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FixedBase(_Handle):
    """Handle of an immutable fixed-base window table (bn254_fixed_base)."""
    _destroy = "bn254_fixed_base_destroy"

    def __init__(self, engine, handle, group):
        self.engine, self.handle, self.group = engine, handle, group


class MsmTable(_Handle):
    """Handle of per-point window tables for shared-point MSMs (bn254_msm_table)."""
    _destroy = "bn254_msm_table_destroy"

    def __init__(self, engine, handle, group, length):
        self.engine, self.handle, self.group, self.len = engine, handle, group, length


def fr_from_ints(vals):
    """ints -> (n, 32) fr.Element (Montgomery form, gnark layout)."""
    return np.frombuffer(b"".join((int(v) % R_MOD * (1 << 256) % R_MOD).to_bytes(32, "little") for v in vals), dtype=np.uint8).reshape(-1, 32).copy()


def fr_to_ints(elems):
    """(n, 32) fr.Element -> ints (regular value)."""
    rinv = pow(1 << 256, -1, R_MOD)
    e = np.ascontiguousarray(elems).reshape(-1, 32)
    return [int.from_bytes(e[i].tobytes(), "little") * rinv % R_MOD for i in range(e.shape[0])]


def fr_lagrange_basis(s, x):
    """Delta_{s_i,S}(x) for every i (utils/compute_lagrange_basis.go:8-30) with one inversion: fr.Element in and out.
    Host-side C++ in the library (no GPU work)."""
    s, x = _u8(s, 32, "s"), _u8(x, 32, "x")
    n = s.size // 32
    out = np.empty(n * 32, dtype=np.uint8)
    fn = _native.lib().bn254_fr_lagrange_basis
    fn.restype = None
    fn(s.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(n), x.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p))
    return out.reshape(n, 32)


def fr_to_scalars(elems):
    """fr.Element array -> regular-form little-endian scalars (x.BigInt(new(big.Int)))."""
    e = _u8(elems, 32, "elems")
    n = e.size // 32
    out = np.empty(n * 32, dtype=np.uint8)
    fn = _native.lib().bn254_fr_to_scalars
    fn.restype = None
    fn(e.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(n), out.ctypes.data_as(ctypes.c_void_p))
    return out.reshape(n, 32)


class G2Lines:
    """Handle of a device-resident line table (bn254_lines)."""

    def __init__(self, engine, handle, m):
        self.engine, self.handle, self.m = engine, handle, m

    def close(self):
        if self.handle:
            fn = self.engine._lib.bn254_g2_lines_destroy
            fn.argtypes = [ctypes.c_void_p]
            fn.restype = None
            fn(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def device_count():
    """Number of CUDA devices the library sees (bn254_device_count); contexts can only be created on B200s."""
    fn = _native.lib().bn254_device_count
    fn.restype = ctypes.c_int
    return int(fn())


_default = None
_default_lock = threading.Lock()


def default_engine():
    global _default
    with _default_lock:
        if _default is None:
            _default = Engine(0)
        return _default


def set_default_engine(e):
    global _default
    _default = e


# =============================================================================================
# gnark-named value types
# =============================================================================================
def _neg_fp_raw(b):
    """Negate one Montgomery-form Fp element given as 32 raw little-endian bytes (p - v, 0 -> 0)."""
    v = int.from_bytes(b, "little")
    return ((P_MOD - v) % P_MOD).to_bytes(32, "little")


def _norm_scalar(s):
    """big.Int semantics for group scalars: returns (negate, s mod r)."""
    s = int(s)
    return (s < 0), abs(s) % R_MOD


class _Point:
    NBYTES = 0
    _mul = _mul_base = _add = None

    def __init__(self, raw=None):
        self.raw = bytes(self.NBYTES) if raw is None else bytes(raw)
        assert len(self.raw) == self.NBYTES

    def Set(self, a):
        self.raw = a.raw
        return self

    def SetInfinity(self):
        self.raw = bytes(self.NBYTES)
        return self

    def IsInfinity(self):
        return self.raw == bytes(self.NBYTES)

    def Equal(self, other):
        return self.raw == other.raw

    def __eq__(self, other):
        return type(self) is type(other) and self.raw == other.raw

    def __hash__(self):
        return hash(self.raw)

    def Neg(self, a):
        half = self.NBYTES // 2
        x, y = a.raw[:half], a.raw[half:]
        self.raw = x + b"".join(_neg_fp_raw(y[i:i + 32]) for i in range(0, half, 32))
        return self

    def ScalarMultiplication(self, a, s):
        neg, k = _norm_scalar(s)
        src = type(self)().Neg(a) if neg else a
        e = default_engine()
        out = getattr(e, self._mul)(src.raw, scalars_to_bytes([k]))
        self.raw = out.tobytes()
        return self

    def Add(self, a, b):
        self.raw = getattr(default_engine(), self._add)(a.raw, b.raw).tobytes()
        return self

    def Sub(self, a, b):
        return self.Add(a, type(self)().Neg(b))


class G1Affine(_Point):
    NBYTES = G1_BYTES
    _mul, _add = "g1_mul_batch", "g1_add_batch"

    def Bytes(self):
        from . import wire
        return wire.g1_bytes(self.raw)

    def Marshal(self):
        from . import wire
        return wire.g1_marshal(self.raw)

    def Unmarshal(self, b):
        from . import wire
        self.raw = wire.g1_unmarshal(bytes(b))
        return self

    def ScalarMultiplicationBase(self, s):
        return self.ScalarMultiplication(Generators()[2], s)


class G2Affine(_Point):
    NBYTES = G2_BYTES
    _mul, _add = "g2_mul_batch", "g2_add_batch"

    def Bytes(self):
        from . import wire
        return wire.g2_bytes(self.raw)

    def Marshal(self):
        from . import wire
        return wire.g2_marshal(self.raw)

    def Unmarshal(self, b):
        from . import wire
        self.raw = wire.g2_unmarshal(bytes(b))
        return self

    def ScalarMultiplicationBase(self, s):
        return self.ScalarMultiplication(Generators()[3], s)


_GT_ONE = None


def _gt_one_raw():
    global _GT_ONE
    if _GT_ONE is None:
        one = (1 << 256) % P_MOD
        _GT_ONE = one.to_bytes(32, "little") + bytes(GT_BYTES - 32)
    return _GT_ONE


class GT:
    """bn254.GT (= fptower.E12).  Zero value is 0, not 1, exactly as in Go."""

    def __init__(self, raw=None):
        self.raw = bytes(GT_BYTES) if raw is None else bytes(raw)
        assert len(self.raw) == GT_BYTES

    def Set(self, a):
        self.raw = a.raw
        return self

    def SetOne(self):
        self.raw = _gt_one_raw()
        return self

    def IsZero(self):
        return self.raw == bytes(GT_BYTES)

    def Equal(self, other):
        return self.raw == other.raw

    def __eq__(self, other):
        return isinstance(other, GT) and self.raw == other.raw

    def __hash__(self):
        return hash(self.raw)

    def Mul(self, a, b):
        self.raw = default_engine().gt_mul_batch(a.raw, b.raw).tobytes()
        return self

    def Div(self, a, b):
        self.raw = default_engine().gt_div_batch(a.raw, b.raw).tobytes()
        return self

    def Inverse(self, a):
        self.raw = default_engine().gt_div_batch(_gt_one_raw(), a.raw).tobytes()
        return self

    def Bytes(self):
        from . import wire
        return wire.gt_bytes(self.raw)

    Marshal = Bytes

    def Unmarshal(self, b):
        from . import wire
        self.raw = wire.gt_from_bytes(bytes(b))
        return self

    def Exp(self, x, k):
        """z = x^k; k == 0 -> 1; k < 0 -> (x^-1)^|k| (gnark E12.Exp semantics)."""
        k = int(k)
        base = x
        if k < 0:
            base = GT().Inverse(x)
            k = -k
        if k >> 256:
            raise ValueError("GT.Exp exponent must be below 2^256")
        self.raw = default_engine().gt_exp_batch(base.raw, scalars_to_bytes([k])).tobytes()
        return self


def Generators():
    """(g1Jac, g2Jac, g1Aff, g2Aff) -- Jacobian forms are returned as (affine, z=1) pairs."""
    g1 = (ctypes.c_uint8 * G1_BYTES)()
    g2 = (ctypes.c_uint8 * G2_BYTES)()
    _native.lib().bn254_generators(g1, g2)
    a1, a2 = G1Affine(bytes(g1)), G2Affine(bytes(g2))
    return (a1, 1), (a2, 1), a1, a2


def HashToG1(msg, dst):
    """bn254.HashToG1(msg, dst []byte) (G1Affine, error)"""
    return G1Affine(default_engine().hash_to_g1_batch([msg], dst)[0].tobytes())


def HashToG2(msg, dst):
    """bn254.HashToG2(msg, dst []byte) (G2Affine, error)"""
    return G2Affine(default_engine().hash_to_g2_batch([msg], dst)[0].tobytes())


def _pack(Ps, Qs):
    if len(Ps) == 0 or len(Ps) != len(Qs):
        raise ValueError("invalid inputs sizes")
    return b"".join(p.raw for p in Ps), b"".join(q.raw for q in Qs), len(Ps)


def Pair(P, Q):
    """bn254.Pair(P []G1Affine, Q []G2Affine) (GT, error)"""
    pb, qb, k = _pack(P, Q)
    return GT(default_engine().multi_pair_batch(pb, qb, k).tobytes())


def PairingCheck(P, Q):
    """bn254.PairingCheck(P, Q) (bool, error)"""
    pb, qb, k = _pack(P, Q)
    return bool(default_engine().pairing_check_batch(pb, qb, k)[0])


def MillerLoop(P, Q):
    """bn254.MillerLoop(P, Q) (GT, error): defined up to factors FinalExponentiation removes."""
    pb, qb, k = _pack(P, Q)
    return GT(default_engine().miller_loop_batch(pb, qb, k).tobytes())


def FinalExponentiation(z, *more):
    """bn254.FinalExponentiation(z *GT, _z ...*GT) GT: the inputs are multiplied first."""
    acc = z
    for m in more:
        acc = GT().Mul(acc, m)
    return GT(default_engine().final_exp_batch(acc.raw).tobytes())
