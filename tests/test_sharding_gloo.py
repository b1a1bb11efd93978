"""N>1 host logic on CPU: world_size-2 gloo processes shard a batch and a big multi-pairing exactly as
bench.py / the Go scheduler do; partials recombine to the unsharded answer (computed with the oracle here,
because there is no GPU in this container)."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from gopairingbasedcryptography_b200.sharding import shard_range


def test_shard_range_covers_exactly():
    for n in (0, 1, 7, 1024, (1 << 20) + 3):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert all(hi - lo <= -(-n // world) for lo, hi in spans)


def _worker(rank, world, port_no, k, q):
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import torch

    from oracle import port

    import common

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P, Q, _, _ = common.points(k, seed=4321, threads=1)
    lo, hi = shard_range(k, rank, world)
    # per-rank partial Miller product over its contiguous share of the k pairs
    part = port.miller_loop_batch(P[64 * lo:64 * hi], Q[128 * lo:128 * hi], 1, hi - lo)
    t = torch.from_numpy(part.copy())
    gathered = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(gathered, t)
    # independent-batch sharding: pairings of my slice
    mine = port.pair_batch(P[64 * lo:64 * hi], Q[128 * lo:128 * hi], hi - lo)
    sizes = [shard_range(k, r, world) for r in range(world)]
    outs = [torch.empty((b - a) * 384, dtype=torch.uint8) for a, b in sizes]
    dist.all_gather(outs, torch.from_numpy(mine.copy())) if len({b - a for a, b in sizes}) == 1 else None
    if rank == 0:
        acc = gathered[0].numpy()
        for g in gathered[1:]:
            acc = port.gt_mul_batch(acc, g.numpy(), 1)
        combined = port.final_exp_batch(acc, 1)
        full = port.multi_pair_batch(P, Q, 1, k)
        batch_ok = True
        if len({b - a for a, b in sizes}) == 1:
            batch_ok = bool((np.concatenate([o_.numpy() for o_ in outs]) == port.pair_batch(P, Q, k)).all())
        q.put((bool((combined == full).all()), batch_ok))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_partition_and_partial_product_combine():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port_no = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port_no, 6, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok == (True, True)
