"""Batch partitioning across GPUs: contiguous chunks, no data-path collective.

Two users:
  * one process per GPU (bench.py under torchrun): `shard_range` picks the rank's slice;
  * ONE process driving every GPU of the box -- the shape of the Go host (go/bn254/engine.go: `shard`): `DevicePool`
    holds one engine context per device and one worker thread per context; every batch entry point splits its batch
    into ceil(n/G) contiguous chunks, runs them concurrently (the C calls release the GIL and each context has its own
    streams and pinned staging) and concatenates the results in order.

The only exchange step north_star names -- combining per-GPU partial Miller products of ONE very large multi-pairing --
is `DevicePool.multi_pair_split` / `combine_partials`: a host-side gather of <= 8 x 384 B followed by GT products and
one final exponentiation on a single GPU."""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor

import numpy as np

G1_BYTES, G2_BYTES, GT_BYTES, SCALAR_BYTES = 64, 128, 384, 32


def shard_range(n, rank, world):
    """Contiguous [lo, hi) owned by `rank`: ceil(n/world) per rank, the tail ranks may be short or empty."""
    per = -(-n // world)
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def shard_pairs_of_product(k, rank, world):
    """Which of the k pairs of one big multi-pairing a rank accumulates."""
    return shard_range(k, rank, world)


def combine_partials(engine, partials):
    """partials: sequence of (384,) uint8 Miller-loop partial products -> final GT (384,) uint8."""
    acc = np.ascontiguousarray(partials[0]).reshape(1, GT_BYTES)
    for p in partials[1:]:
        acc = engine.gt_mul_batch(acc, np.ascontiguousarray(p).reshape(1, GT_BYTES))
    return engine.final_exp_batch(acc)[0]


def _rows(x, item_bytes):
    a = np.frombuffer(x, dtype=np.uint8) if isinstance(x, (bytes, bytearray, memoryview)) else np.ascontiguousarray(x).view(np.uint8).reshape(-1)
    if a.size % item_bytes:
        raise ValueError("invalid inputs sizes")
    return a.reshape(-1, item_bytes)


class DevicePool:
    """One engine context per device, batches split ceil(n/G) per device, all devices in parallel from one process.

    engines: explicit list of bn254.Engine (two contexts on ONE GPU are allowed -- that is how the tests cover the
    dispatcher on a single-GPU box); otherwise one context per ordinal in `devices` (default: every visible GPU).
    min_per_device: batches smaller than this per device use fewer devices (a chunk far below one wave of threads
    only adds launch latency), like `shard(n, minPerDevice, ...)` in the Go package."""

    def __init__(self, devices=None, engines=None, min_per_device=1024):
        from . import bn254

        if engines is None:
            if devices is None:
                devices = range(bn254.device_count())
            engines = [bn254.Engine(d) for d in devices]
            self._owned = True
        else:
            self._owned = False
        if not engines:
            raise bn254.EngineError("no B200 device (this engine has no CPU fallback)")
        self.engines = list(engines)
        self.min_per_device = int(min_per_device)
        self._workers = ThreadPoolExecutor(max_workers=len(self.engines), thread_name_prefix="bn254-dev")

    def close(self):
        self._workers.shutdown(wait=True)
        if self._owned:
            for e in self.engines:
                e.close()
        self.engines = []

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __len__(self):
        return len(self.engines)

    # ---- the dispatcher -----------------------------------------------------------------------------------------
    def spans(self, n, min_per_device=None):
        """[(engine, lo, hi)] for a batch of n items."""
        per = self.min_per_device if min_per_device is None else min_per_device
        world = max(1, min(len(self.engines), -(-n // per)))
        out = []
        for r in range(world):
            lo, hi = shard_range(n, r, world)
            if hi > lo:
                out.append((self.engines[r], lo, hi))
        return out

    def shard(self, n, fn, min_per_device=None):
        """Runs fn(engine, lo, hi) for every chunk, concurrently; returns the chunk results in order.  The first
        failure is re-raised after every chunk has finished (no call is left running on a context)."""
        spans = self.spans(n, min_per_device)
        if len(spans) <= 1:
            return [fn(*s) for s in spans]
        futs = [self._workers.submit(fn, *s) for s in spans]
        res, err = [], None
        for f in futs:
            try:
                res.append(f.result())
            except Exception as e:  # noqa: BLE001 -- collected, re-raised below
                err = err or e
        if err is not None:
            raise err
        return res

    def _cat(self, parts, width, dtype=np.uint8):
        if not parts:
            return np.empty((0, width), dtype=dtype) if width else np.empty(0, dtype=dtype)
        return np.concatenate(parts, axis=0)

    # ---- pairings (the gnark call surface as batch entry points; reference call sites in include/bn254_b200.h) ---
    def pair_batch(self, P, Q, out=None):
        """n independent pairings.  Every device writes its chunk straight into `out` (allocated here when not given;
        page-locked P, Q and out are copied to / from the devices without staging), so nothing is concatenated."""
        P, Q = _rows(P, G1_BYTES), _rows(Q, G2_BYTES)
        if len(P) != len(Q):
            raise ValueError("invalid inputs sizes")
        if out is None:
            out = np.empty((len(P), GT_BYTES), dtype=np.uint8)
        elif not isinstance(out, np.ndarray) or not out.flags["C_CONTIGUOUS"] or out.size * out.itemsize != len(P) * GT_BYTES:
            raise ValueError("invalid inputs sizes")
        rows = out.reshape(-1).view(np.uint8).reshape(len(P), GT_BYTES)
        self.shard(len(P), lambda e, lo, hi: e.pair_batch(P[lo:hi], Q[lo:hi], out=rows[lo:hi]))
        return rows

    def _kpairs(self, method, P, Q, k, width):
        P, Q = _rows(P, G1_BYTES), _rows(Q, G2_BYTES)
        k = int(k)
        if k <= 0 or len(P) != len(Q) or len(P) % k:
            raise ValueError("invalid inputs sizes")
        n = len(P) // k
        return self._cat(self.shard(n, lambda e, lo, hi: getattr(e, method)(P[lo * k:hi * k], Q[lo * k:hi * k], k)), width)

    def multi_pair_batch(self, P, Q, k):
        return self._kpairs("multi_pair_batch", P, Q, k, GT_BYTES)

    def miller_loop_batch(self, P, Q, k=1):
        return self._kpairs("miller_loop_batch", P, Q, k, GT_BYTES)

    def pairing_check_batch(self, P, Q, k):
        return self._kpairs("pairing_check_batch", P, Q, k, 0)

    def pairing_check2_fixed_g1_batch(self, p0, p1, q0, q1):
        q0, q1 = _rows(q0, G2_BYTES), _rows(q1, G2_BYTES)
        if len(q0) != len(q1):
            raise ValueError("invalid inputs sizes")
        return self._cat(self.shard(len(q0), lambda e, lo, hi: e.pairing_check2_fixed_g1_batch(p0, p1, q0[lo:hi], q1[lo:hi])), 0)

    def final_exp_batch(self, f):
        f = _rows(f, GT_BYTES)
        return self._cat(self.shard(len(f), lambda e, lo, hi: e.final_exp_batch(f[lo:hi])), GT_BYTES)

    def multi_pair_split(self, P, Q):
        """ONE multi-pairing of k pairs, its pairs spread over the devices: every device accumulates the Miller
        product of its contiguous share, the <= G partials (384 B each) are multiplied on the first device and go
        through one final exponentiation.  Equals Engine.multi_pair_batch(P, Q, k)[0] bit for bit."""
        P, Q = _rows(P, G1_BYTES), _rows(Q, G2_BYTES)
        k = len(P)
        if k == 0 or k != len(Q):
            raise ValueError("invalid inputs sizes")
        partials = self.shard(k, lambda e, lo, hi: e.miller_loop_batch(P[lo:hi], Q[lo:hi], hi - lo)[0], min_per_device=1)
        return combine_partials(self.engines[0], partials)

    # ---- groups and GT ------------------------------------------------------------------------------------------
    def _mul(self, method, base, scalars, pt_bytes, broadcast):
        s = _rows(scalars, SCALAR_BYTES)
        if broadcast:
            return self._cat(self.shard(len(s), lambda e, lo, hi: getattr(e, method)(base, s[lo:hi])), pt_bytes)
        b = _rows(base, pt_bytes)
        if len(b) != len(s):
            raise ValueError("invalid inputs sizes")
        return self._cat(self.shard(len(s), lambda e, lo, hi: getattr(e, method)(b[lo:hi], s[lo:hi])), pt_bytes)

    def g1_mul_batch(self, base, scalars):
        return self._mul("g1_mul_batch", base, scalars, G1_BYTES, False)

    def g2_mul_batch(self, base, scalars):
        return self._mul("g2_mul_batch", base, scalars, G2_BYTES, False)

    def g1_mul_base_batch(self, base1, scalars):
        return self._mul("g1_mul_base_batch", base1, scalars, G1_BYTES, True)

    def g2_mul_base_batch(self, base1, scalars):
        return self._mul("g2_mul_base_batch", base1, scalars, G2_BYTES, True)

    def gt_exp_batch(self, x, k):
        return self._mul("gt_exp_batch", x, k, GT_BYTES, False)

    def gt_cyclo_exp_batch(self, x, k):
        return self._mul("gt_cyclo_exp_batch", x, k, GT_BYTES, False)

    def gt_cyclo_exp_base_batch(self, x1, k):
        return self._mul("gt_cyclo_exp_base_batch", x1, k, GT_BYTES, True)

    def _binary(self, method, a, b, width):
        a, b = _rows(a, width), _rows(b, width)
        if len(a) != len(b):
            raise ValueError("invalid inputs sizes")
        return self._cat(self.shard(len(a), lambda e, lo, hi: getattr(e, method)(a[lo:hi], b[lo:hi])), width)

    def gt_mul_batch(self, a, b):
        return self._binary("gt_mul_batch", a, b, GT_BYTES)

    def gt_div_batch(self, a, b):
        return self._binary("gt_div_batch", a, b, GT_BYTES)

    def gt_cyclo_div_batch(self, a, b):
        return self._binary("gt_cyclo_div_batch", a, b, GT_BYTES)

    def g1_add_batch(self, a, b):
        return self._binary("g1_add_batch", a, b, G1_BYTES)

    def g2_add_batch(self, a, b):
        return self._binary("g2_add_batch", a, b, G2_BYTES)

    def _hash(self, method, msgs, dst, width):
        msgs = msgs if isinstance(msgs, list) else list(msgs)
        return self._cat(self.shard(len(msgs), lambda e, lo, hi: getattr(e, method)(msgs[lo:hi], dst)), width)

    def hash_to_g1_batch(self, msgs, dst):
        return self._hash("hash_to_g1_batch", msgs, dst, G1_BYTES)

    def hash_to_g2_batch(self, msgs, dst):
        return self._hash("hash_to_g2_batch", msgs, dst, G2_BYTES)
