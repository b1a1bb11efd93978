"""The single-process multi-GPU dispatcher (sharding.DevicePool; the Go host's `shard`, go/bn254/engine.go).
CPU part: the dispatcher's logic over stand-in engines backed by the oracle's C restatement (no GPU here).
GPU part: two contexts of one GPU behave as two devices -- pool results == one engine's, bit for bit; one big
multi-pairing split over the contexts and recombined (north_star's only cross-GPU step) == the unsplit product."""
import threading

import numpy as np
import pytest

from gopairingbasedcryptography_b200.sharding import DevicePool, combine_partials, shard_range
from oracle import port

import common


class OracleEngine:
    """Stand-in with the Engine method names the pool calls, computing with the oracle (test infrastructure)."""

    def __init__(self, fail_on=None):
        self.calls, self.fail_on, self.threads = [], fail_on, set()

    def _note(self, name, n):
        self.calls.append((name, n))
        self.threads.add(threading.get_ident())
        if self.fail_on == name:
            raise RuntimeError("injected failure in %s" % name)

    def pair_batch(self, P, Q, out=None):
        n = P.size // 64
        self._note("pair_batch", n)
        res = port.pair_batch(np.ascontiguousarray(P).reshape(-1), np.ascontiguousarray(Q).reshape(-1), n, 1).reshape(n, 384)
        if out is not None:
            out[...] = res
        return res

    def miller_loop_batch(self, P, Q, k=1):
        n = P.size // 64 // k
        self._note("miller_loop_batch", n)
        return port.miller_loop_batch(np.ascontiguousarray(P).reshape(-1), np.ascontiguousarray(Q).reshape(-1), n, k).reshape(n, 384)

    def multi_pair_batch(self, P, Q, k):
        n = P.size // 64 // k
        self._note("multi_pair_batch", n)
        return port.multi_pair_batch(np.ascontiguousarray(P).reshape(-1), np.ascontiguousarray(Q).reshape(-1), n, k).reshape(n, 384)

    def gt_mul_batch(self, a, b):
        n = a.size // 384
        self._note("gt_mul_batch", n)
        return port.gt_mul_batch(np.ascontiguousarray(a).reshape(-1), np.ascontiguousarray(b).reshape(-1), n).reshape(n, 384)

    def final_exp_batch(self, f):
        n = f.size // 384
        self._note("final_exp_batch", n)
        return port.final_exp_batch(np.ascontiguousarray(f).reshape(-1), n).reshape(n, 384)

    def close(self):
        pass


def test_pool_splits_contiguously_and_keeps_order():
    P, Q, _, _ = common.points(11, seed=77, threads=2)
    engines = [OracleEngine() for _ in range(3)]
    pool = DevicePool(engines=engines, min_per_device=1)
    out = pool.pair_batch(P, Q)
    assert (out.reshape(-1) == port.pair_batch(P, Q, 11, 2)).all()
    assert [e.calls for e in engines] == [[("pair_batch", 4)], [("pair_batch", 4)], [("pair_batch", 3)]]  # ceil(11/3) per device
    assert len(set().union(*[e.threads for e in engines])) >= 2  # chunks ran on worker threads, not one after the other on the caller's
    pool.close()


def test_pool_small_batches_stay_on_fewer_devices_and_empty_batches_work():
    P, Q, _, _ = common.points(5, seed=78, threads=2)
    engines = [OracleEngine() for _ in range(4)]
    pool = DevicePool(engines=engines, min_per_device=4)
    assert [(hi - lo) for _, lo, hi in pool.spans(5)] == [3, 2]
    assert [(hi - lo) for _, lo, hi in pool.spans(3)] == [3]
    assert pool.pair_batch(P[:0], Q[:0]).shape == (0, 384)
    out = pool.multi_pair_batch(P[:64 * 4], Q[:128 * 4], 2)
    assert (out.reshape(-1) == port.multi_pair_batch(P[:64 * 4], Q[:128 * 4], 2, 2)).all()
    with pytest.raises(ValueError, match="invalid inputs sizes"):
        pool.pair_batch(P, Q[:128])
    with pytest.raises(ValueError, match="invalid inputs sizes"):
        pool.multi_pair_batch(P, Q, 2)  # 5 pairs do not split into products of 2
    pool.close()


def test_pool_failure_on_one_device_is_raised_after_all_chunks_finish():
    P, Q, _, _ = common.points(6, seed=79, threads=2)
    engines = [OracleEngine(), OracleEngine(fail_on="pair_batch"), OracleEngine()]
    pool = DevicePool(engines=engines, min_per_device=1)
    with pytest.raises(RuntimeError, match="injected failure"):
        pool.pair_batch(P, Q)
    assert all(len(e.calls) == 1 for e in engines)  # the healthy devices were not abandoned mid-call
    pool.close()


def test_split_multi_pairing_recombines_to_the_unsplit_product():
    k = 7
    P, Q, _, _ = common.points(k, seed=80, threads=2)
    engines = [OracleEngine() for _ in range(3)]
    pool = DevicePool(engines=engines, min_per_device=1000)  # the split ignores min_per_device: it is ONE item
    got = pool.multi_pair_split(P, Q)
    assert (got == port.multi_pair_batch(P, Q, 1, k)).all()
    assert [c for c in engines[1].calls] == [("miller_loop_batch", 1)] and ("final_exp_batch", 1) in engines[0].calls
    spans = [shard_range(k, r, 3) for r in range(3)]
    parts = [port.miller_loop_batch(P[64 * lo:64 * hi], Q[128 * lo:128 * hi], 1, hi - lo) for lo, hi in spans]
    assert (combine_partials(engines[0], parts) == got).all()
    pool.close()


@pytest.mark.gpu
def test_pool_over_two_contexts_of_one_gpu_matches_one_engine(engine):
    from gopairingbasedcryptography_b200 import bn254

    n = 2500
    P, Q, a, b = common.points(64, seed=81)
    P, Q = np.tile(P.reshape(-1, 64), (n // 64 + 1, 1))[:n], np.tile(Q.reshape(-1, 128), (n // 64 + 1, 1))[::-1][:n].copy()
    sc = common.scalar_bytes(common.scalars(n, seed=82)).reshape(n, 32)
    second = bn254.Engine(engine.device)
    with DevicePool(engines=[engine, second], min_per_device=256) as pool:
        assert (pool.pair_batch(P, Q) == engine.pair_batch(P, Q)).all()
        assert (pool.multi_pair_batch(P[:2400], Q[:2400], 3) == engine.multi_pair_batch(P[:2400], Q[:2400], 3)).all()
        assert (pool.pairing_check_batch(P[:2400], Q[:2400], 2) == engine.pairing_check_batch(P[:2400], Q[:2400], 2)).all()
        assert (pool.g1_mul_batch(P, sc) == engine.g1_mul_batch(P, sc)).all()
        assert (pool.g2_mul_batch(Q, sc) == engine.g2_mul_batch(Q, sc)).all()
        assert (pool.g1_mul_base_batch(P[0], sc) == engine.g1_mul_base_batch(P[0], sc)).all()
        gt = engine.pair_batch(P[:600], Q[:600])
        assert (pool.gt_cyclo_exp_batch(gt, sc[:600]) == engine.gt_cyclo_exp_batch(gt, sc[:600])).all()
        assert (pool.gt_mul_batch(gt, gt[::-1].copy()) == engine.gt_mul_batch(gt, gt[::-1].copy())).all()
        msgs = [b"pool message %d" % i for i in range(700)]
        assert (pool.hash_to_g2_batch(msgs, b"DST") == engine.hash_to_g2_batch(msgs, b"DST")).all()
        ok = pool.pairing_check2_fixed_g1_batch(P[0], P[1], Q[:900], Q[900:1800])
        assert (ok == engine.pairing_check2_fixed_g1_batch(P[0], P[1], Q[:900], Q[900:1800])).all()
    second.close()


@pytest.mark.gpu
def test_one_multi_pairing_split_over_two_contexts(engine):
    """sharding.combine_partials on the GPU: 201 pairs (the BSW07-100 product) split 101 + 100 over two contexts."""
    from gopairingbasedcryptography_b200 import bn254

    k = 201
    P, Q, _, _ = common.points(k, seed=83)
    second = bn254.Engine(engine.device)
    with DevicePool(engines=[engine, second]) as pool:
        got = pool.multi_pair_split(P, Q)
    second.close()
    assert (got == engine.multi_pair_batch(P, Q, k)[0]).all()
    assert (got == port.multi_pair_batch(P, Q, 1, k)).all()
