#!/usr/bin/env python
"""Headline benchmark: raw batched bn254.Pair on 2^20 random (G1,G2) pairs per GPU (BASELINE.json
configs[1]).  One "step" = one pass of the hot path over one batch.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--log2-batch 20] [--impl reference]

N > 1 is launched by torchrun (one rank per GPU); the batch is sharded by rank, no collective on the
data path ("weak" scaling: 2^20 pairs per GPU).  Prints ONE JSON line (rank 0).

 value     pairings/s, inputs resident in HBM, CUDA events on the launching stream, max over ranks
 e2e       pairings/s through the public host-buffer API (H2D + kernel + D2H inside the timed region)
 roofline  IMAD (32x32->64 multiply-accumulate) pipe: algorithmic limb-MACs (2.081e6 per
           pairing) / kernel time, against the IMAD.WIDE rate measured live on the same GPU
 cpu_baseline  the oracle's C restatement timed on this box's host cores on a bounded sample
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "bn254_pairings_per_sec"
UNIT = "pairings/s"
MACS_PER_PAIRING = 2.081e6  # SURVEY.md §8d Model-M: 15300 Fp-mul equivalents x 136 limb-MACs
BYTES_PER_PAIRING = 64 + 128 + 384
# dram__bytes_read.sum + dram__bytes_write.sum of k_pair from the committed ncu --set full capture of round 2
# (profiles/r2/ncu_k_pair_r2_summary.txt, 2^18 pairings): what is left of the local-memory stack traffic after the staged
# tower (188.5 GB / 2^18 before it), still ~160x the algorithmic 576 B per pairing.
NCU_DRAM_BYTES_PER_PAIRING = (3.396864e9 + 21.053526e9) / (1 << 18)


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_sample(n, threads):
    """Time the oracle's C restatement (CPU stand-in for gnark, which cannot run here) on n pairings."""
    from oracle import bn254_ref as o
    from oracle import port

    rng = o.SplitMix64(0xB2000254 + 2)
    g1, g2 = port.generators()
    m = 64
    sb = np.frombuffer(b"".join(o.scalar_to_bytes(rng.scalar()) for _ in range(2 * m)), dtype=np.uint8)
    P = port.g1_mul_base_batch(g1, sb[: 32 * m], m, threads)
    Q = port.g2_mul_base_batch(g2, sb[32 * m:], m, threads)
    reps = (n + m - 1) // m
    P, Q = np.tile(P, reps)[: 64 * n], np.tile(Q, reps)[: 128 * n]
    port.pair_batch(P[: 64 * threads], Q[: 128 * threads], threads, threads)  # warm
    t0 = time.perf_counter()
    port.pair_batch(P, Q, n, threads)
    return n / (time.perf_counter() - t0)


def run_reference(args):
    """--impl reference: the reference's CPU path.  gnark-crypto (Go) cannot be built or run in this image
    (no Go toolchain, module not on disk), so this times the oracle's C restatement of it on all host
    threads; each step is a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_threads()
    sample = max(threads * 256, 2048)
    for _ in range(args.warmup):
        cpu_sample(threads * 4, threads)
    t0 = time.perf_counter()
    rates = [cpu_sample(sample, threads) for _ in range(args.steps)]
    dt = time.perf_counter() - t0
    v = sample * args.steps / sum(sample / r for r in rates)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64 limbs (4x64 Montgomery)", "data": "synthetic",
        "config": {"workload": "raw batched bn254.Pair, 2^%d random (G1,G2) pairs per GPU" % args.log2_batch,
                   "batch_per_gpu": 1 << args.log2_batch, "sample_pairings_per_step": sample},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d pairings per step on %d threads; C restatement of gnark's algorithm (gnark itself "
                                   "cannot run: no Go toolchain)" % (sample, threads)},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self.stop_flag = index, [], set(), None, False

    def run(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            }
            while not self.stop_flag:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.1)
        except Exception as e:  # clocks are evidence, not the product
            self.reasons.add("nvml_unavailable:%s" % type(e).__name__)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


def measure_imad_peak(device):
    """Live IMAD.WIDE rate on this GPU: run the microbenchmark binary's carry-chained kernel if it is built."""
    path = os.path.join(ROOT, "profiles", "microbench", "imad_peak")
    rec = os.path.join(ROOT, "profiles", "microbench", "imad_peak_b200.json")
    try:
        import subprocess

        env = dict(os.environ, CUDA_VISIBLE_DEVICES=str(device))
        out = subprocess.run([path, "--quick"], capture_output=True, text=True, timeout=120, env=env).stdout
        j = json.loads(out)
        return j["imad_wide_x_chain"]["ops_per_s"], "measured live (profiles/microbench/imad_peak --quick)"
    except Exception:
        with open(rec) as f:
            j = json.load(f)
        return j["imad_wide_x_chain"]["ops_per_s"], "recorded (profiles/microbench/imad_peak_b200.json)"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--log2-batch", type=int, default=20)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--cpu-sample", type=int, default=0, help="pairings in the cpu_baseline sample (0 = auto)")
    ap.add_argument("--light", action="store_true", help="kernel sweep mode: skip e2e, cpu_baseline and the live IMAD peak")
    ap.add_argument("--rows", default="full", choices=["none", "quick", "full"], help="the rest of the metric (BLS verifies/s, scalar mults/s, ...)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch

    from gopairingbasedcryptography_b200 import bn254

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):
            os.environ.pop("NCCL_DEBUG", None)  # VERSION / WARN print NCCL's banner on stdout; rank 0 prints ONE JSON line
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        cpu_group = dist.new_group(backend="gloo")  # host-side wait for the one-process leg (an NCCL barrier would spin ON the GPUs it measures)
    torch.cuda.set_device(local)
    eng = bn254.Engine(local)
    n = 1 << args.log2_batch

    # ---- synthetic inputs, generated on the GPU: P = A[i % m] + B[i // m], same for Q ----------------
    from benchmarks.rows import SplitMix64, load_executed_imad, measure_rows  # the product arm imports nothing from oracle/

    rng = SplitMix64(0xB2000254 + 2 + 1000 * rank)
    m = 1 << (args.log2_batch // 2)
    mh = n // m
    sc = [rng.scalar(bn254.R_MOD) for _ in range(2 * (m + mh))]
    sb = bn254.scalars_to_bytes(sc)
    g1, g2 = bn254.Generators()[2:]
    A1 = eng.g1_mul_base_batch(g1.raw, sb[:m]); B1 = eng.g1_mul_base_batch(g1.raw, sb[m:m + mh])
    A2 = eng.g2_mul_base_batch(g2.raw, sb[m + mh:2 * m + mh]); B2 = eng.g2_mul_base_batch(g2.raw, sb[2 * m + mh:])
    P = eng.g1_add_batch(np.tile(A1, (mh, 1)), np.repeat(B1, m, axis=0))
    Q = eng.g2_add_batch(np.tile(A2, (mh, 1)), np.repeat(B2, m, axis=0))
    hP = torch.from_numpy(P).pin_memory(); hQ = torch.from_numpy(Q).pin_memory()
    dP, dQ = hP.cuda(non_blocking=True), hQ.cuda(non_blocking=True)
    dO = torch.empty((n, 384), dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step_dev():
        eng.pair_batch_dev(dP.data_ptr(), dQ.data_ptr(), n, dO.data_ptr(), stream)

    for _ in range(args.warmup):
        step_dev()
    barrier()
    sampler = ClockSampler(local); sampler.start()
    launches0 = eng.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_dev()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = eng.launches - launches0
    t = torch.tensor([ms], device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * n * args.steps / (ms_max * 1e-3)
    # the dominant kernel alone, for the roofline: k_pair over the whole batch (BN254_IMPL=thread never routes a
    # remainder to the lane-group kernel, so one step is exactly one k_pair launch)
    os.environ["BN254_IMPL"] = "thread"
    eng_t = bn254.Engine(local)
    del os.environ["BN254_IMPL"]
    eng_t.pair_batch_dev(dP.data_ptr(), dQ.data_ptr(), n, dO.data_ptr(), stream)
    torch.cuda.synchronize()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = eng_t.launches
    k0.record()
    for _ in range(args.steps):
        eng_t.pair_batch_dev(dP.data_ptr(), dQ.data_ptr(), n, dO.data_ptr(), stream)
    k1.record()
    torch.cuda.synchronize()
    assert eng_t.launches - l0 == args.steps
    kernel_ms = k0.elapsed_time(k1) / args.steps
    step_dev()  # leave the default path's output in dO for the verification below
    torch.cuda.synchronize()

    # ---- strong scaling: ONE 2^20 batch split over the N ranks (BASELINE configs[1] wording) ---------------------
    ns = n // world
    step_s = lambda: eng.pair_batch_dev(dP.data_ptr(), dQ.data_ptr(), ns, dO.data_ptr(), stream)
    if world > 1:
        step_s()
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(args.steps):
            step_s()
        s1.record()
        barrier()
        ts_ = torch.tensor([s0.elapsed_time(s1)], device="cuda")
        dist.all_reduce(ts_, op=dist.ReduceOp.MAX)
        strong_ms = float(ts_.item()) / args.steps
        step_dev()
        torch.cuda.synchronize()
    else:
        strong_ms = ms_max / args.steps
    strong = {"workload": "ONE batch of 2^%d pairs split over %d GPU(s): %d pairs per GPU" % (args.log2_batch, world, ns),
              "value": ns * world / (strong_ms * 1e-3), "unit": UNIT, "ms_per_step": strong_ms,
              "efficiency_vs_weak": (ns * world / (strong_ms * 1e-3)) / value,
              "note": "CTAs are scheduled one by one, so a part-filled last wave costs its share only (2^17 pairs: 45.9 ms against 43.1 ideal)"}

    # ---- end to end through the host-buffer API --------------------------------------------------
    hP_np, hQ_np = hP.numpy(), hQ.numpy()
    if args.light:
        sampler.stop_flag = True
        if rank == 0:
            idx = np.array([0, 1, n // 2, n - 1])
            from oracle import port
            ref = port.pair_batch(P[idx].reshape(-1), Q[idx].reshape(-1), len(idx), 4).reshape(len(idx), 384)
            ok = bool((dO.cpu().numpy()[idx] == ref).all())
            print(json.dumps({"variant": os.environ.get("BN254_VARIANT", ""), "value": value, "ms_per_step": ms_max / args.steps,
                              "log2_batch": args.log2_batch, "parity_sample_ok": ok, "clocks": sampler.summary()}))
        return
    eng.pair_batch(hP_np[:4096], hQ_np[:4096])  # warm the copy path
    # the caller's buffers are page-locked (north_star: pinned host buffers): inputs pinned above, result buffer here
    out_host = torch.empty((n, 384), dtype=torch.uint8).pin_memory().numpy()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.pair_batch(hP_np, hQ_np, out=out_host)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    sampler.stop_flag = True; sampler.join(timeout=2)
    t = torch.tensor([e2e_s], device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * n * args.steps / float(t.item())

    # ---- ONE process driving all N GPUs (the Go host's shape): rank 0 opens a context on every GPU of the job and
    # splits the same 2^20-pair host batch ceil(n/N) per device from N host threads (sharding.DevicePool); the other
    # ranks wait at the barrier with their GPUs idle.  Host buffers, copies inside the timed region.
    one_process = None
    if world > 1:
        barrier()
        if rank == 0:
            from gopairingbasedcryptography_b200.sharding import DevicePool

            pool_out = torch.empty((n, 384), dtype=torch.uint8).pin_memory().numpy()
            with DevicePool(devices=range(world)) as pool:
                pool.pair_batch(hP_np, hQ_np, out=pool_out)
                t0 = time.perf_counter()
                for _ in range(args.steps):
                    pool.pair_batch(hP_np, hQ_np, out=pool_out)
                dt = (time.perf_counter() - t0) / args.steps
            assert (pool_out == out_host).all(), "one-process dispatcher and single-GPU path disagree"
            one_process = {"contexts": world, "host_threads": world, "value": n / dt, "unit": UNIT, "ms_per_step": 1e3 * dt,
                           "workload": "the same ONE batch of 2^%d pairs, host buffers, split over %d GPUs by ONE process" % (args.log2_batch, world),
                           "parity": "== the single-GPU result on all %d pairings" % n}
            del pool_out
        dist.barrier(group=cpu_group)  # the other ranks block on the host here: their GPUs stay idle for rank 0's contexts
        barrier()

    # ---- verification (outside the timed regions): sampled bit-exact parity + device == host path ----
    from oracle import port

    torch.cuda.synchronize()
    dev_out = dO.cpu().numpy()
    assert (dev_out == out_host).all(), "device-pointer and host-buffer paths disagree"
    idx = np.unique(np.concatenate([[0, 1, n // 2, n - 2, n - 1], np.random.default_rng(rank).integers(0, n, 59)]))
    ref = port.pair_batch(P[idx].reshape(-1), Q[idx].reshape(-1), len(idx), host_threads()).reshape(len(idx), 384)
    assert (dev_out[idx] == ref).all(), "GPU pairings differ from the oracle"

    peak, peak_how = measure_imad_peak(local) if rank == 0 else (9.24e12, "")
    rows = []
    if args.rows != "none":
        load_executed_imad(ROOT)
        rows = measure_rows(eng, peak, rank=rank, world=world, dist=dist, quick=(args.rows == "quick"), cpu=True)
    if rank == 0:
        achieved = MACS_PER_PAIRING * n / (kernel_ms * 1e-3)
        threads = host_threads()
        sample = args.cpu_sample or max(threads * 1024, 4096)  # ~20 s of CPU time on the C restatement
        cpu_v = cpu_sample(sample, threads)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32 limbs (8x32 Montgomery, IMAD.WIDE)", "data": "synthetic",
            "config": {"workload": "raw batched bn254.Pair, 2^%d random (G1,G2) pairs per GPU" % args.log2_batch,
                       "batch_per_gpu": n, "l2": "inputs+outputs per step = %d MiB > 126 MB L2" % (n * BYTES_PER_PAIRING >> 20),
                       "parity": "%d sampled outputs bit-exact vs oracle; host path == device path on all %d" % (len(idx), n)},
            "clocks": sampler.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n * 192, "d2h_bytes_per_step": n * 384},
            "gpu_launches": launches,
            "strong_2p20": dict(strong, one_process=one_process),
            "rows": rows,
            "roofline": {"bound": "imad", "achieved": achieved / 1e12, "peak": peak / 1e12, "unit": "T limb-MAC/s",
                         "frac": achieved / peak, "traffic": NCU_DRAM_BYTES_PER_PAIRING * n,
                         "kernel": "k_pair", "kernel_ms": kernel_ms,
                         "note": "algorithmic 2.081e6 32x32->64 MACs per pairing (SURVEY 8d) x %d per k_pair launch / %.1f ms (timed alone on a thread-kernel context); "
                                 "peak = IMAD.WIDE rate %s; algorithmic HBM: %.2f GB/s of %.0f measured (not the bound); traffic = DRAM bytes per launch "
                                 "scaled from the ncu capture at 2^18 (local-memory stack spill, see profiles/r2)"
                                 % (n, kernel_ms, peak_how, BYTES_PER_PAIRING * n / (kernel_ms * 1e-3) / 1e9, 6472.1)},
            "cpu_baseline": {"value": cpu_v, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": "%d pairings on %d threads, C restatement of gnark (gnark itself cannot run here: no Go)" % (sample, threads)},
        }
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
