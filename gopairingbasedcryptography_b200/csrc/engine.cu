// libbn254_b200.so -- host runtime behind include/bn254_b200.h.  The kernels live in one translation unit per family
// (k_pairing.cu, k_group.cu, k_gt.cu, k_hash.cu, k_vm.cu, k_fr.cu) and are reached through launch.h.
// sm_100a only; no CPU fallback: every entry point needs a live CUDA device.
//
// Locking: every public entry point takes ctx->mu ONCE and holds it across table lookup / build, scratch growth, the
// launches and (host-buffer entry points) the copies back -- the *_locked helpers below assume the lock is held.
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <set>
#include <string>
#include <vector>

#include "../../include/bn254_b200.h"
#include "launch.h"

namespace L = bn254::launch;
std::atomic<uint64_t> bn254::launch::g_launches{0};

struct Slot {
  cudaStream_t stream = nullptr;
  char* h = nullptr;  // pinned staging
  char* d = nullptr;  // device staging
  void* vm_cold = nullptr;     // cold slot scratch of the lane-group kernels launched on this stream
  void* mp_scratch = nullptr;  // partial Miller values / GT ladder tables of launches on this stream
  size_t mp_scratch_bytes = 0;
  // pending output copy-back (cleared before every return, never carried across calls)
  void* user_out = nullptr;
  size_t out_off = 0, out_bytes = 0;
  bool busy = false;
};

struct FixedTable {  // immutable once built
  int group = 0;     // BN254_GROUP_*
  void* dev = nullptr;
  unsigned char base[BN254_GT_BYTES] = {};
  uint64_t stamp = 0;  // LRU clock of the implicit cache
};
struct bn254_fixed_base { bn254_ctx* ctx; FixedTable t; };
struct bn254_msm_table { bn254_ctx* ctx; int group; size_t len; void* tables; };
struct bn254_lines {
  bn254_ctx* ctx;
  size_t m;
  void* table;     // m x kLinesPerPoint x 3 Fp2 on the device
  uint8_t* qskip;  // m flags: point at infinity
};

struct bn254_ctx {
  int device = 0;
  std::mutex mu;
  std::string err;
  Slot slot[2];
  size_t slot_bytes = 0;
  // lane-group (tower VM) kernels.  BN254_IMPL: "vm" = lane-group kernels always, "thread" = one-thread-per-pairing
  // kernels always, unset = thread kernels, except that launches of at most kVmAutoMax elements take the lane-group
  // kernels (lower latency: three lanes share one pairing)
  int vm_mode = 0;  // 0 auto, 1 lane-group always, 2 never (thread kernels only), 3 warp-VM always
  int sms = 0;
  int vm_blocks_per_sm[3] = {0, 0, 0};
  int wvm_blocks_per_sm[4] = {0, 0, 0, 0};
  bool g2_glv = false;  // BN254_G2_LADDER=glv: 2-dimensional GLV on G2 instead of the 4-dimensional GLS ladder
  size_t wvm_auto_max = 0;  // launches of at most this many items take the warp-VM kernels (one warp per item)
  void* vm_cold_dev = nullptr;  // scratch for the *_dev entry points (launches are serialised by vm_dev_done)
  cudaEvent_t vm_dev_done = nullptr;
  // implicit fixed-base tables of the *_base_batch entry points: small LRU keyed by the base's bytes.  Entries are
  // only evicted under ctx->mu, which the using call holds until its last launch has been enqueued AND completed
  // (host-buffer entry points return synchronously), so a table is never rebuilt under a running kernel.
  std::vector<FixedTable> cache;
  uint64_t clock = 0;
  // device allocations of live table / line handles: released with the context if the caller never destroyed them
  std::set<void*> handle_mem;
};

namespace {

constexpr size_t kSlotBytes = 96u << 20;      // per-slot staging (pinned + device)
constexpr size_t kMaxChunkItems = 1u << 17;   // keeps two chunks in flight for copy/compute overlap
constexpr size_t kVmAutoMax = 16384;
constexpr size_t kFixedMin = 4096;            // from this many scalars on a one-base call builds / uses a window table
constexpr size_t kCacheEntries = 4;           // implicit tables kept per context (BSW07 keygen alternates g2, g2^alpha)
constexpr size_t kG1Jac = 96, kG2Jac = 192;

int fail(bn254_ctx* c, int code, const char* what, cudaError_t e = cudaSuccess) {
  if (c) {
    c->err = what;
    if (e != cudaSuccess) { c->err += ": "; c->err += cudaGetErrorString(e); }
  }
  return code;
}
inline int cu_code(cudaError_t e) { return e == cudaErrorMemoryAllocation ? BN254_ERR_OOM : BN254_ERR_CUDA; }
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(ctx, cu_code(e_), #call, e_); } while (0)

// Small launches are latency-bound: up to wvm_auto_max items run one WARP per item (warp-VM), up to kVmAutoMax three
// lanes per item (lane-group VM), beyond that one thread per item.  BN254_IMPL = thread | vm | wvm forces one path.
inline bool use_wvm(const bn254_ctx* ctx, size_t n) { return ctx->vm_mode == 3 || (ctx->vm_mode == 0 && n <= ctx->wvm_auto_max); }
// Two-pair products on the warp-VM: ONE warp on the two-pair Miller program (1 221 rounds) or TWO warps on the single-pair
// program (803 rounds) plus a product kernel.  Two warps win while they find idle schedulers; from about two warps per
// scheduler on (n products > a quarter of the warp-VM grid) the shared squarings win (1024 BLS checks: 2.60 -> 2.41 ms,
// 64 checks: 1.58 ms on two warps against 1.88 ms on one; profiles/r2/bls_config0_latency.jsonl).
inline bool use_wvm_miller2(const bn254_ctx* ctx, size_t n) { return use_wvm(ctx, n) && (ctx->vm_mode == 3 ? n > 592 : 4 * n > ctx->wvm_auto_max); }
inline bool use_vm(const bn254_ctx* ctx, size_t n) { return ctx->vm_mode == 1 || ctx->vm_mode == 3 || (ctx->vm_mode == 0 && n <= kVmAutoMax); }
inline size_t pt_bytes(int g) { return g == 1 ? BN254_G1_BYTES : BN254_G2_BYTES; }
inline size_t jac_bytes(int g) { return g == 1 ? kG1Jac : kG2Jac; }
inline int slot_index(const bn254_ctx* ctx, cudaStream_t s) { return s == ctx->slot[1].stream ? 1 : 0; }

// Grows a slot's scratch.  Only called between launches of the slot's own stream, after that stream has drained
// what used the old allocation.
cudaError_t ensure_scratch(Slot& sl, size_t need) {
  if (sl.mp_scratch_bytes >= need) return cudaSuccess;
  if (sl.mp_scratch) {
    cudaError_t e = cudaStreamSynchronize(sl.stream);
    if (e != cudaSuccess) return e;
    cudaFree(sl.mp_scratch); sl.mp_scratch = nullptr; sl.mp_scratch_bytes = 0;
  }
  cudaError_t e = cudaMalloc(&sl.mp_scratch, need);
  if (e == cudaSuccess) sl.mp_scratch_bytes = need;
  return e;
}

// Where a launch sequence gets its temporary device memory from:
//  host-buffer entry points -> the slot's persistent scratch (grown on demand, reused by stream order);
//  *_dev entry points       -> stream-ordered allocations on the CALLER's stream (cudaMallocAsync / cudaFreeAsync), so
//                              concurrent calls on different streams never share scratch.
struct Scratch {
  Slot* slot = nullptr;
  cudaStream_t stream = nullptr;
  std::vector<void*> async;
  size_t used = 0;
  cudaError_t reserve(size_t total) { return slot ? ensure_scratch(*slot, total) : cudaSuccess; }
  cudaError_t get(size_t bytes, void** out) {
    bytes = (bytes + 255) & ~size_t(255);
    if (slot) {
      if (used + bytes > slot->mp_scratch_bytes) return cudaErrorMemoryAllocation;  // reserve() was too small: a bug
      *out = static_cast<char*>(slot->mp_scratch) + used;
      used += bytes;
      return cudaSuccess;
    }
    cudaError_t e = cudaMallocAsync(out, bytes, stream);
    if (e == cudaSuccess) async.push_back(*out);
    return e;
  }
  void release() {
    for (void* p : async) cudaFreeAsync(p, stream);
    async.clear();
    used = 0;
  }
};
inline size_t al256(size_t b) { return (b + 255) & ~size_t(255); }

struct Operand { const void* ptr; size_t item_bytes; bool broadcast; };

void drain_slots(bn254_ctx* ctx) {  // error exit: nothing in flight, nothing pending, no user pointer kept
  for (int i = 0; i < 2; i++) {
    Slot& s = ctx->slot[i];
    cudaStreamSynchronize(s.stream);
    s.busy = false; s.user_out = nullptr; s.out_bytes = 0;
  }
}

// Chunked, double-buffered host-buffer driver: up to three input operands (each per-item or broadcast), one output.
// launch(d_in[3], count, d_out, slot) -> cudaError_t.  ctx->mu is held by the caller.
template <typename LF>
int run_host_locked(bn254_ctx* ctx, const Operand* in, int nin, void* out, size_t out_item, size_t n, LF launch) {
  if (n == 0) return BN254_OK;
  if (!out) return fail(ctx, BN254_ERR_BAD_ARG, "null pointer");
  size_t fixed = 1024, per = out_item;
  for (int k = 0; k < nin; k++) {
    if (in[k].item_bytes && !in[k].ptr) return fail(ctx, BN254_ERR_BAD_ARG, "null pointer");
    if (in[k].broadcast) fixed += in[k].item_bytes + 256; else per += in[k].item_bytes;
  }
  if (fixed >= ctx->slot_bytes) return fail(ctx, BN254_ERR_BAD_ARG, "element too large for staging");
  size_t chunk = std::min<size_t>({n, kMaxChunkItems, (ctx->slot_bytes - fixed) / per});
  if (chunk == 0) return fail(ctx, BN254_ERR_BAD_ARG, "element too large for staging");
  // one-thread-per-item kernels: a chunk of whole waves (148 SMs x 3 CTAs x 128 threads) does not pay for a part-filled
  // last wave in every chunk (131 072 = 2.31 waves ran as 3)
  const size_t wave = (size_t)ctx->sms * 3 * L::kBlockThreads;
  if (chunk < n && chunk > wave) chunk -= chunk % wave;
  // Caller buffers in page-locked memory (bn254_host_alloc, cudaHostRegister, torch pin_memory ...) are copied to /
  // from the device directly, chunk by chunk, on the slot's stream; pageable ones go through the pinned staging area.
  auto page_locked = [](const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
  };
  bool in_pinned[3] = {false, false, false};
  for (int k = 0; k < nin; k++) in_pinned[k] = in[k].item_bytes && !in[k].broadcast && page_locked(in[k].ptr);
  const bool out_pinned = page_locked(out);
  auto finish = [&](Slot& s) -> cudaError_t {
    if (!s.busy) return cudaSuccess;
    cudaError_t e = cudaStreamSynchronize(s.stream);
    if (e != cudaSuccess) return e;
    if (s.out_bytes) memcpy(s.user_out, s.h + s.out_off, s.out_bytes);
    s.busy = false; s.user_out = nullptr; s.out_bytes = 0;
    return cudaSuccess;
  };
  auto body = [&]() -> cudaError_t {
    size_t done = 0;
    int ci = 0;
    cudaError_t e;
    while (done < n) {
      size_t c = std::min(chunk, n - done);
      Slot& s = ctx->slot[ci & 1];
      if ((e = finish(s)) != cudaSuccess) return e;
      size_t off = 0;
      const void* d_in[3] = {nullptr, nullptr, nullptr};
      for (int k = 0; k < nin; k++) {
        d_in[k] = s.d + off;
        if (!in[k].item_bytes) continue;
        size_t l = in[k].broadcast ? in[k].item_bytes : in[k].item_bytes * c;
        const char* src = static_cast<const char*>(in[k].ptr) + (in[k].broadcast ? 0 : done * in[k].item_bytes);
        if (in_pinned[k]) e = cudaMemcpyAsync(s.d + off, src, l, cudaMemcpyHostToDevice, s.stream);
        else { memcpy(s.h + off, src, l); e = cudaMemcpyAsync(s.d + off, s.h + off, l, cudaMemcpyHostToDevice, s.stream); }
        if (e != cudaSuccess) return e;
        off = al256(off + l);
      }
      size_t oo = off, lo = out_item * c;
      if ((e = launch(d_in, c, s.d + oo, s)) != cudaSuccess) return e;
      if ((e = cudaGetLastError()) != cudaSuccess) return e;
      s.user_out = static_cast<char*>(out) + done * out_item;
      if (out_pinned) { e = cudaMemcpyAsync(s.user_out, s.d + oo, lo, cudaMemcpyDeviceToHost, s.stream); s.out_bytes = 0; }
      else { e = cudaMemcpyAsync(s.h + oo, s.d + oo, lo, cudaMemcpyDeviceToHost, s.stream); s.out_bytes = lo; }
      if (e != cudaSuccess) return e;
      s.out_off = oo; s.busy = true;
      done += c; ci++;
    }
    for (int i = 0; i < 2; i++) if ((e = finish(ctx->slot[(ci + i) & 1])) != cudaSuccess) return e;
    return cudaSuccess;
  };
  cudaError_t e = body();
  if (e != cudaSuccess) {
    drain_slots(ctx);  // single error exit: both streams idle, no pending copy into the caller's buffer
    cudaGetLastError();
    return fail(ctx, cu_code(e), "host-buffer batch", e);
  }
  return BN254_OK;
}
// public wrapper: argument check, lock, device
template <typename LF>
int run_host(bn254_ctx* ctx, std::initializer_list<Operand> ops, void* out, size_t out_item, size_t n, LF launch) {
  if (!ctx) return BN254_ERR_BAD_ARG;
  Operand in[3];
  int nin = 0;
  for (const Operand& o : ops) in[nin++] = o;
  std::lock_guard<std::mutex> lk(ctx->mu);
  CU(cudaSetDevice(ctx->device));
  return run_host_locked(ctx, in, nin, out, out_item, n, launch);
}

// Device-pointer entry points: enqueue on the caller's stream, never synchronise.
template <typename LF>
int run_dev(bn254_ctx* ctx, size_t n, void* stream, LF launch) {
  if (!ctx) return BN254_ERR_BAD_ARG;
  if (n == 0) return BN254_OK;
  std::lock_guard<std::mutex> lk(ctx->mu);
  CU(cudaSetDevice(ctx->device));
  Scratch sc;
  sc.stream = (cudaStream_t)stream;
  cudaError_t e = launch(sc);
  sc.release();
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) { cudaGetLastError(); return fail(ctx, cu_code(e), "device-pointer batch", e); }
  return BN254_OK;
}
// lane-group kernels on a caller stream share vm_cold_dev: order them across streams with an event
cudaError_t vm_dev(bn254_ctx* ctx, int prog, const void* a, const void* b, size_t n, void* out, cudaStream_t s) {
  cudaError_t e = cudaStreamWaitEvent(s, ctx->vm_dev_done, 0);
  if (e != cudaSuccess) return e;
  L::vm_run(prog, a, b, n, out, ctx->vm_cold_dev, ctx->sms, ctx->vm_blocks_per_sm, s);
  return cudaEventRecord(ctx->vm_dev_done, s);
}
// `cold`: the scratch of the lane-group kernels for this launch site (slot's own, or the shared *_dev one)
cudaError_t vm_any(bn254_ctx* ctx, Scratch& sc, int prog, const void* a, const void* b, size_t n, void* out) {
  if (use_wvm(ctx, n)) {  // warp-VM: state in shared memory only, no scratch to share or order
    L::wvm_run(prog, a, b, n, out, ctx->sms, ctx->wvm_blocks_per_sm, sc.stream);
    return cudaSuccess;
  }
  if (!sc.slot) return vm_dev(ctx, prog, a, b, n, out, sc.stream);
  L::vm_run(prog, a, b, n, out, sc.slot->vm_cold, ctx->sms, ctx->vm_blocks_per_sm, sc.stream);
  return cudaSuccess;
}

// ---- launch sequences shared by the host-buffer and device-pointer entry points ---------------------------------
cudaError_t seq_pair(bn254_ctx* ctx, Scratch& sc, const void* P, const void* Q, size_t n, void* out) {
  if (use_vm(ctx, n)) return vm_any(ctx, sc, L::kVmPair, P, Q, n, out);
  // (A remainder-aware split -- whole waves on the thread kernel, the remainder on the lane-group kernel -- was measured
  // and dropped: CTAs are scheduled one by one, not wave by wave, so 2^17 pairs = 2.31 "waves" take 45.9 ms, not three
  // waves' 53 ms, and the split version took 48.2 ms.  profiles/r2/strong_scaling_split.jsonl)
  L::pair(P, Q, n, out, sc.stream);
  return cudaSuccess;
}
cudaError_t seq_final_exp(bn254_ctx* ctx, Scratch& sc, const void* in, size_t n, void* out) {
  if (use_vm(ctx, n)) return vm_any(ctx, sc, L::kVmFinalExp, in, nullptr, n, out);
  L::final_exp(in, n, out, sc.stream);
  return cudaSuccess;
}
// multi-pairing: single kernel for small k, split + combine for large k.  Products of 2..16 pairs whose total pair
// count is small are latency-bound in the one-thread-per-product kernels, so they go through the lane-group kernels:
// Miller loop per PAIR (n*k lane groups), product of each k values, final exponentiation per product.
cudaError_t seq_multi_pair(bn254_ctx* ctx, Scratch& sc, int mode, const void* P, const void* Q, size_t n, int k, void* out) {
  cudaError_t e;
  if (k == 1 && mode <= 1 && use_vm(ctx, n)) return vm_any(ctx, sc, mode == 0 ? L::kVmMiller : L::kVmPair, P, Q, n, out);
  if (k >= 2 && k <= 2 * L::kMpChunk && ctx->vm_mode != 2 && use_vm(ctx, n * (size_t)k)) {
    size_t pairs = n * (size_t)k;
    if ((e = sc.reserve(al256(pairs * BN254_GT_BYTES) + al256(n * BN254_GT_BYTES))) != cudaSuccess) return e;
    void *ml, *prod_s;
    if ((e = sc.get(pairs * BN254_GT_BYTES, &ml)) != cudaSuccess || (e = sc.get(n * BN254_GT_BYTES, &prod_s)) != cudaSuccess) return e;
    void* prod = mode == 2 ? prod_s : out;
    if (k == 2 && use_wvm_miller2(ctx, n)) {  // one warp per PRODUCT: the two-pair Miller program (shared squarings), no combine
      L::wvm_run(L::kVmMiller2, P, Q, n, prod, ctx->sms, ctx->wvm_blocks_per_sm, sc.stream);
    } else {
      if ((e = vm_any(ctx, sc, L::kVmMiller, P, Q, pairs, ml)) != cudaSuccess) return e;
      L::mp_combine(0, ml, n, k, prod, sc.stream);
    }
    if (mode >= 1 && (e = vm_any(ctx, sc, L::kVmFinalExp, prod, nullptr, n, prod)) != cudaSuccess) return e;
    if (mode == 2) L::gt_is_one(prod, n, static_cast<uint8_t*>(out), sc.stream);
    return cudaSuccess;
  }
  if (k <= 2 * L::kMpChunk) { L::multi_pair(mode, P, Q, n, k, out, sc.stream); return cudaSuccess; }
  int nchunks = (k + L::kMpChunk - 1) / L::kMpChunk;
  size_t need = n * (size_t)nchunks * BN254_GT_BYTES;
  void* partial;
  if ((e = sc.reserve(al256(need))) != cudaSuccess || (e = sc.get(need, &partial)) != cudaSuccess) return e;
  L::mp_partial(P, Q, n, k, nchunks, partial, sc.stream);
  L::mp_combine(mode, partial, n, nchunks, out, sc.stream);
  return cudaSuccess;
}
cudaError_t seq_multi_pair_lines(bn254_ctx*, Scratch& sc, const void* P, const bn254_lines* Lt, size_t n, void* out) {
  const int m = (int)Lt->m;
  const int nchunks = (m + L::kLinesChunk - 1) / L::kLinesChunk;
  size_t need = n * (size_t)nchunks * BN254_GT_BYTES;
  void* partial;
  cudaError_t e;
  if ((e = sc.reserve(al256(need))) != cudaSuccess || (e = sc.get(need, &partial)) != cudaSuccess) return e;
  L::miller_lines(P, Lt->table, Lt->qskip, n, m, nchunks, partial, sc.stream);
  L::mp_combine(1, partial, n, nchunks, out, sc.stream);
  return cudaSuccess;
}
cudaError_t seq_check2_fixed_g1(bn254_ctx* ctx, Scratch& sc, const void* P01, const void* Q0, const void* Q1, size_t n, uint8_t* ok) {
  // small batch: lane-group kernels (see seq_multi_pair).  Crossover measured on the 2-pair check itself
  // (profiles/r2/check2_routing.jsonl): 12 000 checks 12.5 ms against 15.2 ms on the thread kernel, 16 384: 14.3 / 15.3,
  // 20 000: 19.1 / 19.6 -- so the limit is kVmAutoMax CHECKS, not pairs
  if (ctx->vm_mode != 2 && use_vm(ctx, n)) {
    size_t bp = al256(2 * n * BN254_G1_BYTES), bq = al256(2 * n * BN254_G2_BYTES), bm = al256(2 * n * BN254_GT_BYTES), bo = al256(n * BN254_GT_BYTES);
    void *Pp, *Qp, *ml, *prod;
    cudaError_t e;
    if ((e = sc.reserve(bp + bq + bm + bo)) != cudaSuccess) return e;
    if ((e = sc.get(bp, &Pp)) != cudaSuccess || (e = sc.get(bq, &Qp)) != cudaSuccess || (e = sc.get(bm, &ml)) != cudaSuccess ||
        (e = sc.get(bo, &prod)) != cudaSuccess) return e;
    L::pack_check2(P01, Q0, Q1, n, Pp, Qp, sc.stream);
    if (use_wvm_miller2(ctx, n)) {  // one warp per check: two-pair Miller program
      L::wvm_run(L::kVmMiller2, Pp, Qp, n, prod, ctx->sms, ctx->wvm_blocks_per_sm, sc.stream);
    } else {
      if ((e = vm_any(ctx, sc, L::kVmMiller, Pp, Qp, 2 * n, ml)) != cudaSuccess) return e;
      L::mp_combine(0, ml, n, 2, prod, sc.stream);
    }
    if ((e = vm_any(ctx, sc, L::kVmFinalExp, prod, nullptr, n, prod)) != cudaSuccess) return e;
    L::gt_is_one(prod, n, ok, sc.stream);
    return cudaSuccess;
  }
  L::check2_fixed_g1(P01, Q0, Q1, n, ok, sc.stream);
  return cudaSuccess;
}
// GT ladders: each thread needs a contiguous table slice (4 / 16 Fp12) in global memory; launches are capped at one
// full wave so that consecutive waves reuse the same slices (stream order).
cudaError_t seq_gt_exp(bn254_ctx* ctx, Scratch& sc, int cyclo, const void* x, size_t stride, const void* k, size_t n, void* out) {
  const size_t wave = (size_t)L::gt_wave_threads(ctx->sms);
  const size_t per = (cyclo ? L::kGtCycloTable : L::kGtExpTable) * BN254_GT_BYTES;
  // generic ladder on more than a wave: ONE launch with the table on the stack measured 14 % faster than waves in series
  if (!cyclo && n > wave) { L::gt_exp(0, x, stride, k, n, out, nullptr, sc.stream); return cudaSuccess; }
  size_t need = std::min(n, wave) * per;
  void* tab;
  cudaError_t e;
  if ((e = sc.reserve(al256(need))) != cudaSuccess || (e = sc.get(need, &tab)) != cudaSuccess) return e;
  for (size_t off = 0; off < n; off += wave) {
    size_t c = std::min(wave, n - off);
    L::gt_exp(cyclo, static_cast<const char*>(x) + off * stride * BN254_GT_BYTES, stride, static_cast<const char*>(k) + off * BN254_SCALAR_BYTES, c,
              static_cast<char*>(out) + off * BN254_GT_BYTES, tab, sc.stream);
  }
  return cudaSuccess;
}
// G1: 2-dimensional GLV ladder.  G2: 4-dimensional GLS ladder with its per-thread table in a scratch slice; sub-batches
// keep the scratch under 1 GiB.  BN254_G2_LADDER=glv keeps the 2-dimensional ladder (A/B measurements).
cudaError_t seq_scalar_mul(bn254_ctx* ctx, Scratch& sc, int g, const void* base, size_t stride, const void* scalars, size_t n, void* out) {
  if (g == 1 || ctx->g2_glv) { L::scalar_mul(g, base, stride, scalars, n, out, sc.stream); return cudaSuccess; }
  const size_t per = L::g2_gls_scratch_bytes(1);
  const size_t sub = std::max<size_t>(1, std::min<size_t>(n, (size_t(1) << 30) / per));
  void* scratch;
  cudaError_t e;
  const size_t slices = sub + L::kBlockThreads;  // every thread of the last CTA owns a slice (the CTA shares its inversions)
  if ((e = sc.reserve(al256(slices * per))) != cudaSuccess || (e = sc.get(slices * per, &scratch)) != cudaSuccess) return e;
  for (size_t off = 0; off < n; off += sub) {
    size_t c = std::min(sub, n - off);
    L::scalar_mul_g2_gls(static_cast<const char*>(base) + off * stride * BN254_G2_BYTES, stride, static_cast<const char*>(scalars) + off * BN254_SCALAR_BYTES, c,
                         static_cast<char*>(out) + off * BN254_G2_BYTES, scratch, sc.stream);
  }
  return cudaSuccess;
}
// Waters hash: plain subset sum for small batches, byte-window tables (built per call, 8192 short sums) from 2048 selectors on
constexpr size_t kSubsetTabMin = 2048;
cudaError_t seq_subset_sum(bn254_ctx*, Scratch& sc, int g, const void* U, int m, const uint8_t* sel, size_t n, void* out) {
  if (n < kSubsetTabMin || m < 16) { L::subset_sum(g, U, m, sel, n, out, sc.stream); return cudaSuccess; }
  size_t need = (size_t)((m + 7) / 8) * 256 * pt_bytes(g);
  void* table;
  cudaError_t e;
  if ((e = sc.reserve(al256(need))) != cudaSuccess || (e = sc.get(need, &table)) != cudaSuccess) return e;
  L::subset_sum_tab(g, U, m, sel, n, out, table, sc.stream);
  return cudaSuccess;
}
// out[g] = sum of `len` consecutive affine points per group: passes of 32-way partial sums
cudaError_t seq_segment_sum(bn254_ctx*, Scratch& sc, int g, const void* pts, size_t groups, size_t len, void* out) {
  const int chunk = 32;
  const size_t B = pt_bytes(g);
  size_t n1 = (len + chunk - 1) / chunk;
  if (n1 == 1) { L::segment_sum(g, pts, groups, (int)len, chunk, out, sc.stream); return cudaSuccess; }
  size_t n2 = (n1 + chunk - 1) / chunk;
  void *a, *b = nullptr;
  cudaError_t e;
  if ((e = sc.reserve(al256(groups * n1 * B) + al256(groups * n2 * B))) != cudaSuccess) return e;
  if ((e = sc.get(groups * n1 * B, &a)) != cudaSuccess || (e = sc.get(groups * n2 * B, &b)) != cudaSuccess) return e;
  const void* cur = pts;
  size_t cur_len = len;
  void* bufs[2] = {a, b};
  for (int pass = 0;; pass++) {
    size_t nch = (cur_len + chunk - 1) / chunk;
    void* dst = nch == 1 ? out : bufs[pass & 1];
    L::segment_sum(g, cur, groups, (int)cur_len, chunk, dst, sc.stream);
    if (nch == 1) break;
    cur = dst; cur_len = nch;
  }
  return cudaSuccess;
}
// shared-point MSM over per-point window tables: partial Jacobian sums per (vector, chunk of points), then a tree
cudaError_t seq_msm(bn254_ctx* ctx, Scratch& sc, const bn254_msm_table* T, const void* scalars, size_t nvec, void* out) {
  const int g = T->group;
  const size_t len = T->len, JB = jac_bytes(g);
  // about one wave of threads: chunk = points per thread
  size_t wave = (size_t)ctx->sms * (g == 1 ? 4 : 3) * L::kBlockThreads;
  size_t chunk = std::max<size_t>(1, (nvec * len + wave - 1) / wave);
  if (chunk > len) chunk = len;
  size_t n1 = (len + chunk - 1) / chunk;
  const int fan = 16;
  size_t n2 = (n1 + fan - 1) / fan;
  void *a, *b;
  cudaError_t e;
  if ((e = sc.reserve(al256(nvec * n1 * JB) + al256(nvec * n2 * JB))) != cudaSuccess) return e;
  if ((e = sc.get(nvec * n1 * JB, &a)) != cudaSuccess || (e = sc.get(nvec * n2 * JB, &b)) != cudaSuccess) return e;
  void* bufs[2] = {a, b};
  size_t cur_len = n1;
  if (n1 % L::kBlockThreads == 0) {
    // every CTA lies inside one vector: its 128 sums are added by a shared-memory tree (7 dependent additions) and the
    // CTA writes one partial -- two 16-way serial passes less (AFP25: 1024 partials per vector -> 8)
    L::msm_partial_tree(g, T->tables, scalars, nvec, len, (int)chunk, a, sc.stream);
    cur_len = n1 / L::kBlockThreads;
    if (cur_len > 1 && cur_len <= (size_t)L::kBlockThreads && (cur_len & (cur_len - 1)) == 0) {
      L::jac_tree(g, a, nvec, (int)cur_len, out, sc.stream);
      return cudaSuccess;
    }
  } else {
    L::msm_partial(g, T->tables, scalars, nvec, len, (int)chunk, a, sc.stream);
  }
  for (int pass = 0;; pass++) {
    size_t nch = (cur_len + fan - 1) / fan;
    if (nch == 1) { L::jac_sum(g, bufs[pass & 1], nvec, (int)cur_len, fan, nullptr, out, sc.stream); break; }
    L::jac_sum(g, bufs[pass & 1], nvec, (int)cur_len, fan, bufs[(pass + 1) & 1], nullptr, sc.stream);
    cur_len = nch;
  }
  return cudaSuccess;
}

// ---- fixed-base tables ------------------------------------------------------------------------------------------
// table[w * 255 + d - 1] = [d << 8w] base (G1 / G2, affine) or base^(d << 8w) (GT).  Built with the variable-base
// kernels on the slot-0 stream, synchronously; ctx->mu held.
int build_fixed_table(bn254_ctx* ctx, int group, const void* base, FixedTable* t) {
  const size_t entries = (size_t)L::kFixedWindows * L::kFixedEntries;
  const size_t item = group == BN254_GROUP_GT ? BN254_GT_BYTES : pt_bytes(group);
  void* dev = nullptr;
  CU(cudaMalloc(&dev, entries * item));
  Slot& s = ctx->slot[0];
  unsigned char* h = reinterpret_cast<unsigned char*>(s.h);
  memcpy(h, base, item);
  unsigned char* hs = h + 512;
  memset(hs, 0, entries * 32);
  for (int w = 0; w < L::kFixedWindows; w++)
    for (int d = 1; d <= L::kFixedEntries; d++) hs[((size_t)w * L::kFixedEntries + d - 1) * 32 + w] = (unsigned char)d;
  cudaError_t e = cudaMemcpyAsync(s.d, s.h, 512 + entries * 32, cudaMemcpyHostToDevice, s.stream);
  if (e == cudaSuccess) {
    if (group == BN254_GROUP_GT) L::gt_exp(0, s.d, 0, s.d + 512, entries, dev, nullptr, s.stream);
    else L::scalar_mul(group, s.d, 0, s.d + 512, entries, dev, s.stream);
      e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(s.stream);
  if (e != cudaSuccess) { cudaFree(dev); cudaGetLastError(); return fail(ctx, cu_code(e), "fixed-base table build", e); }
  t->group = group; t->dev = dev;
  memcpy(t->base, base, item);
  return BN254_OK;
}
// implicit cache of the *_base_batch entry points (ctx->mu held)
int cached_table(bn254_ctx* ctx, int group, const void* base, const void** table) {
  const size_t item = group == BN254_GROUP_GT ? BN254_GT_BYTES : pt_bytes(group);
  for (FixedTable& t : ctx->cache)
    if (t.group == group && memcmp(t.base, base, item) == 0) { t.stamp = ++ctx->clock; *table = t.dev; return BN254_OK; }
  if (ctx->cache.size() >= kCacheEntries) {
    // evict the least recently used entry; every stream of this context is idle between calls (host entry points
    // return synchronously under the lock we hold), so no kernel still reads it
    auto lru = std::min_element(ctx->cache.begin(), ctx->cache.end(), [](const FixedTable& a, const FixedTable& b) { return a.stamp < b.stamp; });
    cudaFree(lru->dev);
    ctx->cache.erase(lru);
  }
  FixedTable t;
  int rc = build_fixed_table(ctx, group, base, &t);
  if (rc) return rc;
  t.stamp = ++ctx->clock;
  ctx->cache.push_back(t);
  *table = t.dev;
  return BN254_OK;
}

// Small one-base batches: a table that is ALREADY cached is used whatever n (32 mixed additions instead of a GLV ladder:
// a 1-element ScalarMultiplicationBase 1.38 -> ~0.5 ms), and the generators' tables are built on their first use --
// gnark's ScalarMultiplicationBase always multiplies the generator, so that is the base the reference's 1-element calls
// pass (bls_signature.go:45, waters05_ibe.go:224, bsw07_cpabe.go:69,157).  ctx->mu held.
bool small_batch_uses_table(bn254_ctx* ctx, int group, const void* base) {
  if (group != BN254_GROUP_G1 && group != BN254_GROUP_G2) return false;
  const size_t item = pt_bytes(group);
  for (const FixedTable& t : ctx->cache)
    if (t.group == group && memcmp(t.base, base, item) == 0) return true;
  unsigned char g1[BN254_G1_BYTES], g2[BN254_G2_BYTES];
  bn254_generators(g1, g2);
  return memcmp(base, group == BN254_GROUP_G1 ? g1 : g2, item) == 0;
}

// n messages (concatenated bytes + n+1 offsets) -> points.  Chunks are sized to one staging slot; two slots in
// flight: chunk i's kernel runs while chunk i-1's points are copied back and handed to the caller.
int hash_to_curve_host(bn254_ctx* ctx, int G, const uint8_t* msgs, const uint64_t* offsets, size_t n, const uint8_t* dst, size_t dst_len, void* out) {
  if (!ctx) return BN254_ERR_BAD_ARG;
  if (dst_len > 255) return fail(ctx, BN254_ERR_BAD_ARG, "hash-to-curve: domain separation tag longer than 255 bytes");
  if (n == 0) return BN254_OK;
  if (!msgs && offsets && offsets[n] != offsets[0]) return fail(ctx, BN254_ERR_BAD_ARG, "null pointer");
  if (!offsets || !out || (dst_len && !dst)) return fail(ctx, BN254_ERR_BAD_ARG, "null pointer");
  const size_t out_item = pt_bytes(G);
  std::lock_guard<std::mutex> lk(ctx->mu);
  CU(cudaSetDevice(ctx->device));
  // chunks are capped so that a large batch splits into >= 4 of them (a whole wave of CTAs each at the least)
  const size_t cap = std::min<size_t>(kMaxChunkItems, std::max<size_t>((n + 3) / 4, (size_t)148 * 3 * L::kBlockThreads));
  struct Pending { size_t done = 0, c = 0, o_out = 0; bool live = false; } pend[2];
  auto drain = [&](int i) -> cudaError_t {
    if (!pend[i].live) return cudaSuccess;
    cudaError_t e = cudaStreamSynchronize(ctx->slot[i].stream);
    if (e != cudaSuccess) return e;
    memcpy(static_cast<char*>(out) + pend[i].done * out_item, ctx->slot[i].h + pend[i].o_out, out_item * pend[i].c);
    pend[i].live = false;
    return cudaSuccess;
  };
  bool too_large = false;
  auto body = [&]() -> cudaError_t {
    size_t done = 0;
    cudaError_t e;
    for (int it = 0; done < n; it ^= 1) {
      if ((e = drain(it)) != cudaSuccess) return e;
      Slot& s = ctx->slot[it];
      // largest c with  256 (dst) + 8 (c + 1) + bytes + out_item * c  <=  slot
      size_t c = 0, bytes = 0;
      while (done + c < n && c < cap) {
        size_t len = (size_t)(offsets[done + c + 1] - offsets[done + c]);
        if (512 + 8 * (c + 2) + bytes + len + out_item * (c + 1) + 512 > ctx->slot_bytes) break;
        bytes += len; c++;
      }
      if (c == 0) { too_large = true; return cudaErrorInvalidValue; }
      unsigned char* h = reinterpret_cast<unsigned char*>(s.h);
      memset(h, 0, 256);
      if (dst_len) memcpy(h, dst, dst_len);
      uint64_t* ho = reinterpret_cast<uint64_t*>(h + 256);
      for (size_t j = 0; j <= c; j++) ho[j] = offsets[done + j] - offsets[done];
      size_t o_msg = al256(256 + 8 * (c + 1));
      if (bytes) memcpy(h + o_msg, msgs + offsets[done], bytes);
      size_t o_out = al256(o_msg + bytes);
      if ((e = cudaMemcpyAsync(s.d, s.h, o_out, cudaMemcpyHostToDevice, s.stream)) != cudaSuccess) return e;
      const uint8_t* d = reinterpret_cast<const uint8_t*>(s.d);
      L::hash_to_curve(G, d + o_msg, reinterpret_cast<const uint64_t*>(d + 256), c, d, (uint32_t)dst_len, s.d + o_out, s.stream);
      if ((e = cudaGetLastError()) != cudaSuccess) return e;
      if ((e = cudaMemcpyAsync(s.h + o_out, s.d + o_out, out_item * c, cudaMemcpyDeviceToHost, s.stream)) != cudaSuccess) return e;
      pend[it].done = done; pend[it].c = c; pend[it].o_out = o_out; pend[it].live = true;
      done += c;
    }
    if ((e = drain(0)) != cudaSuccess) return e;
    return drain(1);
  };
  cudaError_t e = body();
  if (e != cudaSuccess) {
    drain_slots(ctx);
    cudaGetLastError();
    if (too_large) return fail(ctx, BN254_ERR_BAD_ARG, "message too large for staging");
    return fail(ctx, cu_code(e), "hash-to-curve batch", e);
  }
  return BN254_OK;
}

}  // namespace

// Live contexts.  Table / line handles point at their context; destroying a handle AFTER its context (a garbage-collected
// host language gives no destruction order) must not touch the freed context: the handle's device memory went with the
// context's device allocations, so only the host struct is released.
static std::mutex g_live_mu;
static std::set<bn254_ctx*> g_live;
// Runs `release` under the context lock when the handle's context is still alive; returns false when it is gone.
template <typename F>
static bool with_live_ctx(bn254_ctx* ctx, F release) {
  std::lock_guard<std::mutex> reg(g_live_mu);
  if (!g_live.count(ctx)) return false;
  std::lock_guard<std::mutex> lk(ctx->mu);
  cudaSetDevice(ctx->device);
  release();
  return true;
}

static void release_handle_mem(bn254_ctx* ctx, void* p) {  // under ctx->mu
  if (p && ctx->handle_mem.erase(p)) cudaFree(p);
}

extern "C" {

int bn254_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int bn254_ctx_create(int device, bn254_ctx** out) {
  if (!out) return BN254_ERR_BAD_ARG;
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || device < 0 || device >= n) return BN254_ERR_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return BN254_ERR_CUDA;
  if (prop.major != 10) return BN254_ERR_CUDA;  // sm_100a binary only
  bn254_ctx* ctx = new bn254_ctx();
  ctx->device = device;
  ctx->slot_bytes = kSlotBytes;
  if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return BN254_ERR_CUDA; }
  for (int i = 0; i < 2; i++) {
    Slot& s = ctx->slot[i];
    if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaHostAlloc((void**)&s.h, kSlotBytes, cudaHostAllocDefault) != cudaSuccess ||
        cudaMalloc((void**)&s.d, kSlotBytes) != cudaSuccess) {
      bn254_ctx_destroy(ctx);
      return BN254_ERR_OOM;
    }
  }
  if (L::pairing_init() != cudaSuccess || L::group_init() != cudaSuccess || L::gt_init() != cudaSuccess || L::hash_init() != cudaSuccess) {
    bn254_ctx_destroy(ctx);
    return BN254_ERR_CUDA;
  }
  const char* impl = getenv("BN254_IMPL");
  ctx->vm_mode = !impl ? 0 : (std::string(impl) == "vm" ? 1 : (std::string(impl) == "thread" ? 2 : (std::string(impl) == "wvm" ? 3 : 0)));
  { const char* gl = getenv("BN254_G2_LADDER"); ctx->g2_glv = gl && std::string(gl) == "glv"; }
  ctx->sms = prop.multiProcessorCount;
  if (L::vm_prepare(ctx->vm_blocks_per_sm) != cudaSuccess || L::wvm_prepare(ctx->wvm_blocks_per_sm) != cudaSuccess) { bn254_ctx_destroy(ctx); return BN254_ERR_CUDA; }
  {  // crossover to the lane-group kernels: one pass of the warp-VM grid, 148 SMs x 4 CTAs x 4 warps = 2368 items
     // (profiles/r2/latency_vs_batch.jsonl: 2048 items 4.2 ms vs 5.8 ms, 3072 items 6.1 ms vs 5.8 ms)
    const char* wm = getenv("BN254_WVM_MAX");
    ctx->wvm_auto_max = wm ? (size_t)atol(wm) : (size_t)ctx->sms * ctx->wvm_blocks_per_sm[L::kVmPair] * L::wvm_items_per_cta();
  }
  size_t cb = L::vm_cold_bytes(ctx->sms, ctx->vm_blocks_per_sm);
  if (cudaMalloc(&ctx->vm_cold_dev, cb) != cudaSuccess || cudaMalloc(&ctx->slot[0].vm_cold, cb) != cudaSuccess ||
      cudaMalloc(&ctx->slot[1].vm_cold, cb) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->vm_dev_done, cudaEventDisableTiming) != cudaSuccess) { bn254_ctx_destroy(ctx); return BN254_ERR_OOM; }
  {  // keep stream-ordered scratch of the *_dev entry points cached in the device's default pool
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      uint64_t keep = UINT64_MAX;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
  }
  { std::lock_guard<std::mutex> reg(g_live_mu); g_live.insert(ctx); }
  *out = ctx;
  return BN254_OK;
}

void bn254_ctx_destroy(bn254_ctx* ctx) {
  if (!ctx) return;
  { std::lock_guard<std::mutex> reg(g_live_mu); g_live.erase(ctx); }  // no-op for a half-built context
  cudaSetDevice(ctx->device);
  for (int i = 0; i < 2; i++) {
    Slot& s = ctx->slot[i];
    if (s.stream) { cudaStreamSynchronize(s.stream); cudaStreamDestroy(s.stream); }
    if (s.h) cudaFreeHost(s.h);
    if (s.d) cudaFree(s.d);
    if (s.vm_cold) cudaFree(s.vm_cold);
    if (s.mp_scratch) cudaFree(s.mp_scratch);
  }
  if (ctx->vm_cold_dev) cudaFree(ctx->vm_cold_dev);
  for (FixedTable& t : ctx->cache) cudaFree(t.dev);
  for (void* p : ctx->handle_mem) cudaFree(p);
  if (ctx->vm_dev_done) cudaEventDestroy(ctx->vm_dev_done);
  delete ctx;
}

const char* bn254_last_error(bn254_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
uint64_t bn254_launch_count(bn254_ctx*) { return L::g_launches.load(); }
int bn254_sm_count(bn254_ctx* ctx) { return ctx ? ctx->sms : 0; }

void* bn254_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
  return p;
}
void bn254_host_free(void* p) { if (p) cudaFreeHost(p); }

// ---- device memory and streams for callers without a CUDA binding of their own (the Go package) ----------------
int bn254_dev_alloc(bn254_ctx* ctx, size_t bytes, void** d_out) {
  if (!ctx || !d_out) return BN254_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  CU(cudaSetDevice(ctx->device));
  CU(cudaMalloc(d_out, bytes ? bytes : 1));
  return BN254_OK;
}
int bn254_dev_free(bn254_ctx* ctx, void* d) {
  if (!ctx) return BN254_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  CU(cudaSetDevice(ctx->device));
  CU(cudaFree(d));
  return BN254_OK;
}
int bn254_dev_upload(bn254_ctx* ctx, void* d_dst, const void* h_src, size_t bytes, void* stream) {
  if (!ctx || (bytes && (!d_dst || !h_src))) return BN254_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  CU(cudaSetDevice(ctx->device));
  CU(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  return BN254_OK;
}
int bn254_dev_download(bn254_ctx* ctx, void* h_dst, const void* d_src, size_t bytes, void* stream) {
  if (!ctx || (bytes && (!h_dst || !d_src))) return BN254_ERR_BAD_ARG;
  {
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  }
  CU(cudaStreamSynchronize((cudaStream_t)stream));  // the bytes are in h_dst at return
  return BN254_OK;
}
int bn254_stream_create(bn254_ctx* ctx, void** stream_out) {
  if (!ctx || !stream_out) return BN254_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  CU(cudaSetDevice(ctx->device));
  cudaStream_t s;
  CU(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  *stream_out = s;
  return BN254_OK;
}
int bn254_stream_destroy(bn254_ctx* ctx, void* stream) {
  if (!ctx) return BN254_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  CU(cudaSetDevice(ctx->device));
  CU(cudaStreamDestroy((cudaStream_t)stream));
  return BN254_OK;
}
int bn254_stream_sync(bn254_ctx* ctx, void* stream) {
  if (!ctx) return BN254_ERR_BAD_ARG;
  CU(cudaStreamSynchronize((cudaStream_t)stream));
  return BN254_OK;
}

void bn254_generators(void* g1, void* g2) {
  static const uint32_t G1[16] = {
#include "generators_g1.inc"
  };
  static const uint32_t G2[32] = {
#include "generators_g2.inc"
  };
  memcpy(g1, G1, 64);
  memcpy(g2, G2, 128);
}

// ---- pairings -------------------------------------------------------------------------------
int bn254_pair_batch_dev(bn254_ctx* ctx, const void* dP, const void* dQ, size_t n, void* d_out, void* stream) {
  return run_dev(ctx, n, stream, [&](Scratch& sc) { return seq_pair(ctx, sc, dP, dQ, n, d_out); });
}
int bn254_pair_batch(bn254_ctx* ctx, const void* P, const void* Q, size_t n, void* out) {
  return run_host(ctx, {{P, BN254_G1_BYTES, false}, {Q, BN254_G2_BYTES, false}}, out, BN254_GT_BYTES, n,
                  [ctx](const void* const* d, size_t c, void* o, Slot& sl) {
                    Scratch sc; sc.slot = &sl; sc.stream = sl.stream;
                    return seq_pair(ctx, sc, d[0], d[1], c, o);
                  });
}
#define MULTI_PAIR_ENTRY(name, MODE, OUT_BYTES, OUT_T)                                                                     \
  int name##_dev(bn254_ctx* ctx, const void* dP, const void* dQ, size_t n, size_t k, OUT_T* d_out, void* stream) {         \
    if (k == 0 || k > (1u << 20)) return fail(ctx, BN254_ERR_INVALID_SIZES, "invalid inputs sizes");                       \
    return run_dev(ctx, n, stream, [&](Scratch& sc) { return seq_multi_pair(ctx, sc, MODE, dP, dQ, n, (int)k, d_out); });  \
  }                                                                                                                        \
  int name(bn254_ctx* ctx, const void* P, const void* Q, size_t n, size_t k, OUT_T* out) {                                 \
    if (k == 0 || k > (1u << 20)) return fail(ctx, BN254_ERR_INVALID_SIZES, "invalid inputs sizes");                       \
    int kk = (int)k;                                                                                                       \
    return run_host(ctx, {{P, BN254_G1_BYTES * k, false}, {Q, BN254_G2_BYTES * k, false}}, out, OUT_BYTES, n,              \
                    [kk, ctx](const void* const* d, size_t c, void* o, Slot& sl) {                                         \
                      Scratch sc; sc.slot = &sl; sc.stream = sl.stream;                                                    \
                      return seq_multi_pair(ctx, sc, MODE, d[0], d[1], c, kk, o);                                          \
                    });                                                                                                    \
  }
MULTI_PAIR_ENTRY(bn254_miller_loop_batch, 0, BN254_GT_BYTES, void)
MULTI_PAIR_ENTRY(bn254_multi_pair_batch, 1, BN254_GT_BYTES, void)
MULTI_PAIR_ENTRY(bn254_pairing_check_batch, 2, 1, uint8_t)

// ---- line tables --------------------------------------------------------------------------------------
int bn254_g2_lines_create(bn254_ctx* ctx, const void* Q, size_t m, bn254_lines** out) {
  if (!ctx || !Q || !out || m == 0) return fail(ctx, BN254_ERR_BAD_ARG, "bad lines arguments");
  std::lock_guard<std::mutex> lk(ctx->mu);
  CU(cudaSetDevice(ctx->device));
  if (m * BN254_G2_BYTES > ctx->slot_bytes) return fail(ctx, BN254_ERR_BAD_ARG, "too many points for one table");
  bn254_lines* Lt = new bn254_lines{ctx, m, nullptr, nullptr};
  size_t bytes = m * (size_t)L::kLinesPerPoint * L::kLineBytes;
  if (cudaMalloc(&Lt->table, bytes) != cudaSuccess || cudaMalloc((void**)&Lt->qskip, m) != cudaSuccess) {
    if (Lt->table) cudaFree(Lt->table);
    delete Lt;
    cudaGetLastError();
    return fail(ctx, BN254_ERR_OOM, "line table allocation");
  }
  Slot& s = ctx->slot[0];
  memcpy(s.h, Q, m * BN254_G2_BYTES);
  cudaError_t e = cudaMemcpyAsync(s.d, s.h, m * BN254_G2_BYTES, cudaMemcpyHostToDevice, s.stream);
  if (e == cudaSuccess) { L::g2_lines(s.d, m, Lt->table, Lt->qskip, s.stream); e = cudaGetLastError(); }
  if (e == cudaSuccess) e = cudaStreamSynchronize(s.stream);
  if (e != cudaSuccess) { cudaFree(Lt->table); cudaFree(Lt->qskip); delete Lt; cudaGetLastError(); return fail(ctx, cu_code(e), "line table build", e); }
  ctx->handle_mem.insert(Lt->table); ctx->handle_mem.insert(Lt->qskip);
  *out = Lt;
  return BN254_OK;
}
void bn254_g2_lines_destroy(bn254_lines* Lt) {
  if (!Lt) return;
  with_live_ctx(Lt->ctx, [&] { release_handle_mem(Lt->ctx, Lt->table); release_handle_mem(Lt->ctx, Lt->qskip); });
  delete Lt;
}
size_t bn254_g2_lines_count(const bn254_lines* Lt) { return Lt ? Lt->m : 0; }
// out[i] = FinalExponentiation(prod_j Miller(P[i*m + j], Q_j)) for the m table points: n products of m pairs
int bn254_multi_pair_lines_batch(bn254_ctx* ctx, const void* P, const bn254_lines* Lt, size_t n, void* out) {
  if (!ctx || !Lt || Lt->ctx != ctx) return fail(ctx, BN254_ERR_BAD_ARG, "line table belongs to another context");
  return run_host(ctx, {{P, BN254_G1_BYTES * Lt->m, false}}, out, BN254_GT_BYTES, n,
                  [=](const void* const* d, size_t c, void* o, Slot& sl) {
                    Scratch sc; sc.slot = &sl; sc.stream = sl.stream;
                    return seq_multi_pair_lines(ctx, sc, d[0], Lt, c, o);
                  });
}
int bn254_multi_pair_lines_batch_dev(bn254_ctx* ctx, const void* dP, const bn254_lines* Lt, size_t n, void* d_out, void* stream) {
  if (!ctx || !Lt || Lt->ctx != ctx) return fail(ctx, BN254_ERR_BAD_ARG, "line table belongs to another context");
  return run_dev(ctx, n, stream, [&](Scratch& sc) { return seq_multi_pair_lines(ctx, sc, dP, Lt, n, d_out); });
}

int bn254_final_exp_batch_dev(bn254_ctx* ctx, const void* d_in, size_t n, void* d_out, void* stream) {
  return run_dev(ctx, n, stream, [&](Scratch& sc) { return seq_final_exp(ctx, sc, d_in, n, d_out); });
}
int bn254_final_exp_batch(bn254_ctx* ctx, const void* in, size_t n, void* out) {
  return run_host(ctx, {{in, BN254_GT_BYTES, false}}, out, BN254_GT_BYTES, n,
                  [ctx](const void* const* d, size_t c, void* o, Slot& sl) {
                    Scratch sc; sc.slot = &sl; sc.stream = sl.stream;
                    return seq_final_exp(ctx, sc, d[0], c, o);
                  });
}

// ---- scalar multiplication --------------------------------------------------------------------
#define MUL_ENTRY(name, G, BYTES)                                                                                     \
  int name(bn254_ctx* ctx, const void* base, const void* scalars, size_t n, void* out) {                              \
    return run_host(ctx, {{base, BYTES, false}, {scalars, BN254_SCALAR_BYTES, false}}, out, BYTES, n,                 \
                    [ctx](const void* const* d, size_t c, void* o, Slot& sl) {                                        \
                      Scratch sc; sc.slot = &sl; sc.stream = sl.stream;                                               \
                      return seq_scalar_mul(ctx, sc, G, d[0], 1, d[1], c, o);                                         \
                    });                                                                                               \
  }                                                                                                                   \
  int name##_dev(bn254_ctx* ctx, const void* d_base, size_t stride, const void* d_s, size_t n, void* d_out, void* stream) { \
    return run_dev(ctx, n, stream, [&](Scratch& sc) { return seq_scalar_mul(ctx, sc, G, d_base, stride, d_s, n, d_out); }); \
  }
MUL_ENTRY(bn254_g1_mul_batch, 1, BN254_G1_BYTES)
MUL_ENTRY(bn254_g2_mul_batch, 2, BN254_G2_BYTES)

// One base, n scalars.  Small batches run the GLV kernel on the broadcast base; from kFixedMin scalars on a
// 32 x 255 affine window table of the base is looked up in / added to the context's cache (built once with 8160
// GLV multiplications of d << 8w) and every scalar costs 32 mixed additions.  The lock is held from the lookup to
// the completion of the last launch, so a concurrent caller with another base cannot touch the table in between.
#define MUL_BASE_ENTRY(name, G, BYTES)                                                                                \
  int name(bn254_ctx* ctx, const void* base, const void* scalars, size_t n, void* out) {                              \
    if (!ctx || !base) return BN254_ERR_BAD_ARG;                                                                      \
    std::lock_guard<std::mutex> lk(ctx->mu);                                                                          \
    CU(cudaSetDevice(ctx->device));                                                                                   \
    Operand in[2] = {{base, BYTES, true}, {scalars, BN254_SCALAR_BYTES, false}};                                      \
    if (n < kFixedMin && !small_batch_uses_table(ctx, G, base))                                                       \
      return run_host_locked(ctx, in, 2, out, BYTES, n, [ctx](const void* const* d, size_t c, void* o, Slot& sl) {    \
        Scratch sc; sc.slot = &sl; sc.stream = sl.stream;                                                             \
        return seq_scalar_mul(ctx, sc, G, d[0], 0, d[1], c, o);                                                       \
      });                                                                                                             \
    const void* table = nullptr;                                                                                      \
    int rc = cached_table(ctx, G, base, &table);                                                                      \
    if (rc) return rc;                                                                                                \
    return run_host_locked(ctx, in, 2, out, BYTES, n, [table](const void* const* d, size_t c, void* o, Slot& sl) {    \
      L::fixed_mul(G, table, d[1], c, o, sl.stream);                                                                  \
      return cudaSuccess;                                                                                             \
    });                                                                                                               \
  }
MUL_BASE_ENTRY(bn254_g1_mul_base_batch, 1, BN254_G1_BYTES)
MUL_BASE_ENTRY(bn254_g2_mul_base_batch, 2, BN254_G2_BYTES)

// ---- explicit fixed-base handles (immutable tables owned by the caller) -----------------------------------------
int bn254_fixed_base_create(bn254_ctx* ctx, int group, const void* base, bn254_fixed_base** out) {
  if (!ctx || !base || !out || group < BN254_GROUP_G1 || group > BN254_GROUP_GT) return fail(ctx, BN254_ERR_BAD_ARG, "bad fixed-base arguments");
  std::lock_guard<std::mutex> lk(ctx->mu);
  CU(cudaSetDevice(ctx->device));
  bn254_fixed_base* h = new bn254_fixed_base{ctx, FixedTable()};
  int rc = build_fixed_table(ctx, group, base, &h->t);
  if (rc) { delete h; return rc; }
  ctx->handle_mem.insert(h->t.dev);
  *out = h;
  return BN254_OK;
}
void bn254_fixed_base_destroy(bn254_fixed_base* h) {
  if (!h) return;
  with_live_ctx(h->ctx, [&] { release_handle_mem(h->ctx, h->t.dev); });
  delete h;
}
int bn254_fixed_base_group(const bn254_fixed_base* h) { return h ? h->t.group : 0; }
#define FIXED_ENTRY(name, G, BYTES, LAUNCH)                                                                           \
  int name(bn254_ctx* ctx, const bn254_fixed_base* h, const void* scalars, size_t n, void* out) {                     \
    if (!ctx || !h || h->ctx != ctx || h->t.group != G) return fail(ctx, BN254_ERR_BAD_ARG, "fixed-base handle of another context or group"); \
    const void* table = h->t.dev;                                                                                     \
    return run_host(ctx, {{scalars, BN254_SCALAR_BYTES, false}}, out, BYTES, n,                                       \
                    [table](const void* const* d, size_t c, void* o, Slot& sl) { LAUNCH(table, d[0], c, o, sl.stream); return cudaSuccess; }); \
  }                                                                                                                   \
  int name##_dev(bn254_ctx* ctx, const bn254_fixed_base* h, const void* d_scalars, size_t n, void* d_out, void* stream) { \
    if (!ctx || !h || h->ctx != ctx || h->t.group != G) return fail(ctx, BN254_ERR_BAD_ARG, "fixed-base handle of another context or group"); \
    const void* table = h->t.dev;                                                                                     \
    return run_dev(ctx, n, stream, [&](Scratch& sc) { LAUNCH(table, d_scalars, n, d_out, sc.stream); return cudaSuccess; }); \
  }
#define LAUNCH_FIXED_G1(t, k, c, o, s) L::fixed_mul(1, t, k, c, o, s)
#define LAUNCH_FIXED_G2(t, k, c, o, s) L::fixed_mul(2, t, k, c, o, s)
#define LAUNCH_FIXED_GT(t, k, c, o, s) L::gt_fixed_exp(t, k, c, o, s)
FIXED_ENTRY(bn254_g1_fixed_mul_batch, BN254_GROUP_G1, BN254_G1_BYTES, LAUNCH_FIXED_G1)
FIXED_ENTRY(bn254_g2_fixed_mul_batch, BN254_GROUP_G2, BN254_G2_BYTES, LAUNCH_FIXED_G2)
FIXED_ENTRY(bn254_gt_fixed_exp_batch, BN254_GROUP_GT, BN254_GT_BYTES, LAUNCH_FIXED_GT)

// ---- shared-point MSM -------------------------------------------------------------------------------------------
int bn254_msm_table_create(bn254_ctx* ctx, int group, const void* points, size_t len, bn254_msm_table** out) {
  if (!ctx || !points || !out || len == 0 || (group != BN254_GROUP_G1 && group != BN254_GROUP_G2)) return fail(ctx, BN254_ERR_BAD_ARG, "bad MSM table arguments");
  std::lock_guard<std::mutex> lk(ctx->mu);
  CU(cudaSetDevice(ctx->device));
  const size_t B = pt_bytes(group), F = B / 2;
  if (len * B > ctx->slot_bytes) return fail(ctx, BN254_ERR_BAD_ARG, "too many points for one table");
  const size_t entries = len * (size_t)L::kMsmWindows * L::kFixedEntries;
  void *tables = nullptr, *zs = nullptr, *pf = nullptr;
  cudaError_t e = cudaMalloc(&tables, entries * B);
  if (e == cudaSuccess) e = cudaMalloc(&zs, entries * F);
  if (e == cudaSuccess) e = cudaMalloc(&pf, entries * F);
  Slot& s = ctx->slot[0];
  if (e == cudaSuccess) {
    memcpy(s.h, points, len * B);
    e = cudaMemcpyAsync(s.d, s.h, len * B, cudaMemcpyHostToDevice, s.stream);
  }
  if (e == cudaSuccess) { L::msm_tables(group, s.d, len, tables, zs, pf, s.stream); e = cudaGetLastError(); }
  if (e == cudaSuccess) e = cudaStreamSynchronize(s.stream);
  cudaFree(zs); cudaFree(pf);
  if (e != cudaSuccess) { cudaFree(tables); cudaGetLastError(); return fail(ctx, cu_code(e), "MSM table build", e); }
  ctx->handle_mem.insert(tables);
  *out = new bn254_msm_table{ctx, group, len, tables};
  return BN254_OK;
}
void bn254_msm_table_destroy(bn254_msm_table* T) {
  if (!T) return;
  with_live_ctx(T->ctx, [&] { release_handle_mem(T->ctx, T->tables); });
  delete T;
}
size_t bn254_msm_table_len(const bn254_msm_table* T) { return T ? T->len : 0; }
int bn254_msm_batch(bn254_ctx* ctx, const bn254_msm_table* T, const void* scalars, size_t nvec, void* out) {
  if (!ctx || !T || T->ctx != ctx) return fail(ctx, BN254_ERR_BAD_ARG, "MSM table of another context");
  return run_host(ctx, {{scalars, BN254_SCALAR_BYTES * T->len, false}}, out, pt_bytes(T->group), nvec,
                  [=](const void* const* d, size_t c, void* o, Slot& sl) {
                    Scratch sc; sc.slot = &sl; sc.stream = sl.stream;
                    return seq_msm(ctx, sc, T, d[0], c, o);
                  });
}
int bn254_msm_batch_dev(bn254_ctx* ctx, const bn254_msm_table* T, const void* d_scalars, size_t nvec, void* d_out, void* stream) {
  if (!ctx || !T || T->ctx != ctx) return fail(ctx, BN254_ERR_BAD_ARG, "MSM table of another context");
  return run_dev(ctx, nvec, stream, [&](Scratch& sc) { return seq_msm(ctx, sc, T, d_scalars, nvec, d_out); });
}

// ---- subset sums and segment sums -----------------------------------------------------------------
#define SUBSET_SUM_ENTRY(name, G, BYTES)                                                                                 \
  int name(bn254_ctx* ctx, const void* U, size_t m, const void* sel, size_t n, void* out) {                              \
    if (!ctx || !U || m == 0 || (m + 1) * BYTES > L::subset_sum_max_bytes()) return fail(ctx, BN254_ERR_BAD_ARG, "bad subset-sum arguments"); \
    int mm = (int)m;                                                                                                     \
    return run_host(ctx, {{U, (m + 1) * BYTES, true}, {sel, (m + 7) / 8, false}}, out, BYTES, n,                         \
                    [mm, ctx](const void* const* d, size_t c, void* o, Slot& sl) {                                       \
                      Scratch sc; sc.slot = &sl; sc.stream = sl.stream;                                                  \
                      return seq_subset_sum(ctx, sc, G, d[0], mm, static_cast<const uint8_t*>(d[1]), c, o);              \
                    });                                                                                                  \
  }                                                                                                                      \
  int name##_dev(bn254_ctx* ctx, const void* dU, size_t m, const void* d_sel, size_t n, void* d_out, void* stream) {     \
    if (!ctx || !dU || m == 0 || (m + 1) * BYTES > L::subset_sum_max_bytes()) return fail(ctx, BN254_ERR_BAD_ARG, "bad subset-sum arguments"); \
    return run_dev(ctx, n, stream, [&](Scratch& sc) {                                                                    \
      return seq_subset_sum(ctx, sc, G, dU, (int)m, static_cast<const uint8_t*>(d_sel), n, d_out);                       \
    });                                                                                                                  \
  }
SUBSET_SUM_ENTRY(bn254_g1_subset_sum_batch, 1, BN254_G1_BYTES)
SUBSET_SUM_ENTRY(bn254_g2_subset_sum_batch, 2, BN254_G2_BYTES)

// out[g] = sum of points[g*len .. g*len+len): passes of 32-way partial sums, all on the device
#define SEGMENT_SUM_ENTRY(name, G, BYTES)                                                                                \
  int name(bn254_ctx* ctx, const void* pts, size_t groups, size_t len, void* out) {                                      \
    if (!ctx || !pts || !out || len == 0 || len > (1u << 24)) return fail(ctx, BN254_ERR_BAD_ARG, "bad segment-sum arguments"); \
    if (len * BYTES + 4096 > kSlotBytes) return fail(ctx, BN254_ERR_BAD_ARG, "segment too long for staging");            \
    return run_host(ctx, {{pts, len * BYTES, false}}, out, BYTES, groups,                                                \
                    [=](const void* const* d, size_t c, void* o, Slot& sl) {                                             \
                      Scratch sc; sc.slot = &sl; sc.stream = sl.stream;                                                  \
                      return seq_segment_sum(ctx, sc, G, d[0], c, len, o);                                               \
                    });                                                                                                  \
  }                                                                                                                      \
  int name##_dev(bn254_ctx* ctx, const void* d_pts, size_t groups, size_t len, void* d_out, void* stream) {              \
    if (!ctx || len == 0 || len > (1u << 24)) return fail(ctx, BN254_ERR_BAD_ARG, "bad segment-sum arguments");          \
    return run_dev(ctx, groups, stream, [&](Scratch& sc) { return seq_segment_sum(ctx, sc, G, d_pts, groups, len, d_out); }); \
  }
SEGMENT_SUM_ENTRY(bn254_g1_sum_batch, 1, BN254_G1_BYTES)
SEGMENT_SUM_ENTRY(bn254_g2_sum_batch, 2, BN254_G2_BYTES)

#define ADD_ENTRY(name, G, BYTES)                                                                                        \
  int name(bn254_ctx* ctx, const void* a, const void* b, size_t n, void* out) {                                          \
    return run_host(ctx, {{a, BYTES, false}, {b, BYTES, false}}, out, BYTES, n,                                          \
                    [](const void* const* d, size_t c, void* o, Slot& sl) { L::aff_add(G, d[0], d[1], c, o, sl.stream); return cudaSuccess; }); \
  }                                                                                                                      \
  int name##_dev(bn254_ctx* ctx, const void* da, const void* db, size_t n, void* d_out, void* stream) {                  \
    return run_dev(ctx, n, stream, [&](Scratch& sc) { L::aff_add(G, da, db, n, d_out, sc.stream); return cudaSuccess; }); \
  }
ADD_ENTRY(bn254_g1_add_batch, 1, BN254_G1_BYTES)
ADD_ENTRY(bn254_g2_add_batch, 2, BN254_G2_BYTES)
int bn254_g1_neg_batch_dev(bn254_ctx* ctx, const void* d_in, size_t n, void* d_out, void* stream) {
  return run_dev(ctx, n, stream, [&](Scratch& sc) { L::neg_points(1, d_in, n, d_out, sc.stream); return cudaSuccess; });
}
int bn254_g2_neg_batch_dev(bn254_ctx* ctx, const void* d_in, size_t n, void* d_out, void* stream) {
  return run_dev(ctx, n, stream, [&](Scratch& sc) { L::neg_points(2, d_in, n, d_out, sc.stream); return cudaSuccess; });
}

// ---- GT ---------------------------------------------------------------------------------------
int bn254_gt_exp_batch_dev(bn254_ctx* ctx, const void* d_x, size_t stride, const void* d_k, size_t n, void* d_out, void* stream) {
  return run_dev(ctx, n, stream, [&](Scratch& sc) { return seq_gt_exp(ctx, sc, 0, d_x, stride, d_k, n, d_out); });
}
int bn254_gt_cyclo_exp_batch_dev(bn254_ctx* ctx, const void* d_x, size_t stride, const void* d_k, size_t n, void* d_out, void* stream) {
  return run_dev(ctx, n, stream, [&](Scratch& sc) { return seq_gt_exp(ctx, sc, 1, d_x, stride, d_k, n, d_out); });
}
#define GT_EXP_ENTRY(name, CYCLO)                                                                                        \
  int name(bn254_ctx* ctx, const void* x, const void* k, size_t n, void* out) {                                          \
    return run_host(ctx, {{x, BN254_GT_BYTES, false}, {k, BN254_SCALAR_BYTES, false}}, out, BN254_GT_BYTES, n,           \
                    [ctx](const void* const* d, size_t c, void* o, Slot& sl) {                                           \
                      Scratch sc; sc.slot = &sl; sc.stream = sl.stream;                                                  \
                      return seq_gt_exp(ctx, sc, CYCLO, d[0], 1, d[1], c, o);                                            \
                    });                                                                                                  \
  }
GT_EXP_ENTRY(bn254_gt_exp_batch, 0)
GT_EXP_ENTRY(bn254_gt_cyclo_exp_batch, 1)
#define GT_EXP_BASE_ENTRY(name, CYCLO)                                                                                   \
  int name(bn254_ctx* ctx, const void* x1, const void* k, size_t n, void* out) {                                         \
    if (!ctx || !x1) return BN254_ERR_BAD_ARG;                                                                           \
    std::lock_guard<std::mutex> lk(ctx->mu);                                                                             \
    CU(cudaSetDevice(ctx->device));                                                                                      \
    Operand in[2] = {{x1, BN254_GT_BYTES, true}, {k, BN254_SCALAR_BYTES, false}};                                        \
    if (n < kFixedMin)                                                                                                   \
      return run_host_locked(ctx, in, 2, out, BN254_GT_BYTES, n, [ctx](const void* const* d, size_t c, void* o, Slot& sl) { \
        Scratch sc; sc.slot = &sl; sc.stream = sl.stream;                                                                \
        return seq_gt_exp(ctx, sc, CYCLO, d[0], 0, d[1], c, o);                                                          \
      });                                                                                                                \
    const void* table = nullptr;                                                                                         \
    int rc = cached_table(ctx, BN254_GROUP_GT, x1, &table);                                                              \
    if (rc) return rc;                                                                                                   \
    return run_host_locked(ctx, in, 2, out, BN254_GT_BYTES, n, [table](const void* const* d, size_t c, void* o, Slot& sl) { \
      L::gt_fixed_exp(table, d[1], c, o, sl.stream);                                                                     \
      return cudaSuccess;                                                                                                \
    });                                                                                                                  \
  }
GT_EXP_BASE_ENTRY(bn254_gt_exp_base_batch, 0)
GT_EXP_BASE_ENTRY(bn254_gt_cyclo_exp_base_batch, 1)
#define GT_MUL_ENTRY(name, MODE)                                                                                         \
  int name(bn254_ctx* ctx, const void* a, const void* b, size_t n, void* out) {                                          \
    return run_host(ctx, {{a, BN254_GT_BYTES, false}, {b, BN254_GT_BYTES, false}}, out, BN254_GT_BYTES, n,               \
                    [](const void* const* d, size_t c, void* o, Slot& sl) { L::gt_mul(MODE, d[0], 1, d[1], 1, c, o, sl.stream); return cudaSuccess; }); \
  }                                                                                                                      \
  int name##_dev(bn254_ctx* ctx, const void* da, size_t a_stride, const void* db, size_t b_stride, size_t n, void* d_out, void* stream) { \
    return run_dev(ctx, n, stream, [&](Scratch& sc) { L::gt_mul(MODE, da, a_stride, db, b_stride, n, d_out, sc.stream); return cudaSuccess; }); \
  }
GT_MUL_ENTRY(bn254_gt_mul_batch, 0)
GT_MUL_ENTRY(bn254_gt_div_batch, 1)
GT_MUL_ENTRY(bn254_gt_cyclo_div_batch, 2)
int bn254_fp_mul_batch(bn254_ctx* ctx, const void* a, const void* b, size_t n, void* out) {
  return run_host(ctx, {{a, 32, false}, {b, 32, false}}, out, 32, n,
                  [](const void* const* d, size_t c, void* o, Slot& sl) { L::fp_mul(d[0], d[1], c, o, sl.stream); return cudaSuccess; });
}

// ---- BLS-shaped check: two G1 points fixed for the batch ---------------------------------------------------------
int bn254_pairing_check2_fixed_g1_batch(bn254_ctx* ctx, const void* P01, const void* Q0, const void* Q1, size_t n, uint8_t* ok) {
  if (n && (!P01 || !Q0 || !Q1)) return fail(ctx, BN254_ERR_BAD_ARG, "null pointer");
  return run_host(ctx, {{P01, 2 * BN254_G1_BYTES, true}, {Q0, BN254_G2_BYTES, false}, {Q1, BN254_G2_BYTES, false}}, ok, 1, n,
                  [ctx](const void* const* d, size_t c, void* o, Slot& sl) {
                    Scratch sc; sc.slot = &sl; sc.stream = sl.stream;
                    return seq_check2_fixed_g1(ctx, sc, d[0], d[1], d[2], c, static_cast<uint8_t*>(o));
                  });
}
int bn254_pairing_check2_fixed_g1_batch_dev(bn254_ctx* ctx, const void* dP01, const void* dQ0, const void* dQ1, size_t n, uint8_t* d_ok, void* stream) {
  return run_dev(ctx, n, stream, [&](Scratch& sc) { return seq_check2_fixed_g1(ctx, sc, dP01, dQ0, dQ1, n, d_ok); });
}

// ---- Fr feeders ---------------------------------------------------------------------------------------------------
int bn254_fr_poly_from_roots_dev(bn254_ctx* ctx, const void* d_roots, size_t n, void* d_coeffs, void* stream) {
  if (n == 0 || n > 3000) return fail(ctx, BN254_ERR_BAD_ARG, "polynomial degree out of range (1..3000)");
  return run_dev(ctx, n, stream, [&](Scratch& sc) { return L::fr_poly_from_roots(d_roots, n, d_coeffs, sc.stream); });
}
int bn254_fr_quotient_coeffs_dev(bn254_ctx* ctx, const void* d_f, size_t n, const void* d_ids, size_t nvec, void* d_out, void* stream) {
  if (n == 0 || n > (1u << 24)) return fail(ctx, BN254_ERR_BAD_ARG, "polynomial degree out of range");
  return run_dev(ctx, nvec, stream, [&](Scratch& sc) { L::fr_quotient_coeffs(d_f, n, d_ids, nvec, d_out, sc.stream); return cudaSuccess; });
}
int bn254_fr_to_scalars_dev(bn254_ctx* ctx, const void* d_in, size_t n, void* d_out, void* stream) {
  return run_dev(ctx, n, stream, [&](Scratch& sc) { L::fr_to_scalars(d_in, n, d_out, sc.stream); return cudaSuccess; });
}
int bn254_fr_poly_from_roots(bn254_ctx* ctx, const void* roots, size_t n, void* coeffs) {
  if (!ctx || !roots || !coeffs) return fail(ctx, BN254_ERR_BAD_ARG, "null pointer");
  if (n == 0 || n > 3000) return fail(ctx, BN254_ERR_BAD_ARG, "polynomial degree out of range (1..3000)");
  std::lock_guard<std::mutex> lk(ctx->mu);
  CU(cudaSetDevice(ctx->device));
  Slot& s = ctx->slot[0];
  size_t off = al256(n * 32);
  memcpy(s.h, roots, n * 32);
  cudaError_t e = cudaMemcpyAsync(s.d, s.h, n * 32, cudaMemcpyHostToDevice, s.stream);
  if (e == cudaSuccess) e = L::fr_poly_from_roots(s.d, n, s.d + off, s.stream);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(s.h + off, s.d + off, (n + 1) * 32, cudaMemcpyDeviceToHost, s.stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s.stream);
  if (e != cudaSuccess) { drain_slots(ctx); cudaGetLastError(); return fail(ctx, cu_code(e), "fr_poly_from_roots", e); }
  memcpy(coeffs, s.h + off, (n + 1) * 32);
  return BN254_OK;
}
int bn254_fr_quotient_coeffs(bn254_ctx* ctx, const void* f, size_t n, const void* ids, size_t nvec, void* out) {
  if (!ctx || !f) return fail(ctx, BN254_ERR_BAD_ARG, "null pointer");
  if (n == 0 || (n + 1) * 32 + 4096 > kSlotBytes / 2) return fail(ctx, BN254_ERR_BAD_ARG, "polynomial degree out of range");
  return run_host(ctx, {{f, (n + 1) * 32, true}, {ids, 32, false}}, out, n * 32, nvec,
                  [n](const void* const* d, size_t c, void* o, Slot& sl) { L::fr_quotient_coeffs(d[0], n, d[1], c, o, sl.stream); return cudaSuccess; });
}
void bn254_fr_lagrange_basis(const void* s, size_t n, const void* x, void* out) { L::fr_lagrange_basis_host(s, n, x, out); }
void bn254_fr_to_scalars(const void* in, size_t n, void* out) { L::fr_to_scalars_host(in, n, out); }

// ---- hash-to-curve (gnark bn254.HashToG1 / HashToG2) -----------------------------------------------------------
int bn254_hash_to_g1_batch(bn254_ctx* ctx, const uint8_t* msgs, const uint64_t* offsets, size_t n, const uint8_t* dst, size_t dst_len, void* out) {
  return hash_to_curve_host(ctx, 1, msgs, offsets, n, dst, dst_len, out);
}
int bn254_hash_to_g2_batch(bn254_ctx* ctx, const uint8_t* msgs, const uint64_t* offsets, size_t n, const uint8_t* dst, size_t dst_len, void* out) {
  return hash_to_curve_host(ctx, 2, msgs, offsets, n, dst, dst_len, out);
}
static int hash_dev(bn254_ctx* ctx, int G, const uint8_t* d_msgs, const uint64_t* d_offsets, size_t n, const uint8_t* dst, size_t dst_len, void* d_out, void* stream) {
  if (dst_len > 255) return fail(ctx, BN254_ERR_BAD_ARG, "hash-to-curve: domain separation tag longer than 255 bytes");
  if (n && (!d_offsets || !d_out || (dst_len && !dst))) return fail(ctx, BN254_ERR_BAD_ARG, "null pointer");
  return run_dev(ctx, n, stream, [&](Scratch& sc) {
    void* d_dst;
    cudaError_t e = sc.get(256, &d_dst);
    if (e != cudaSuccess) return e;
    unsigned char tag[256] = {0};
    if (dst_len) memcpy(tag, dst, dst_len);
    // pageable source: the runtime stages the 256 bytes before returning, so `tag` may go out of scope
    if ((e = cudaMemcpyAsync(d_dst, tag, 256, cudaMemcpyHostToDevice, sc.stream)) != cudaSuccess) return e;
    L::hash_to_curve(G, d_msgs, d_offsets, n, static_cast<const uint8_t*>(d_dst), (uint32_t)dst_len, d_out, sc.stream);
    return cudaSuccess;
  });
}
int bn254_hash_to_g1_batch_dev(bn254_ctx* ctx, const uint8_t* d_msgs, const uint64_t* d_offsets, size_t n, const uint8_t* dst, size_t dst_len, void* d_out, void* stream) {
  return hash_dev(ctx, 1, d_msgs, d_offsets, n, dst, dst_len, d_out, stream);
}
int bn254_hash_to_g2_batch_dev(bn254_ctx* ctx, const uint8_t* d_msgs, const uint64_t* d_offsets, size_t n, const uint8_t* dst, size_t dst_len, void* d_out, void* stream) {
  return hash_dev(ctx, 2, d_msgs, d_offsets, n, dst, dst_len, d_out, stream);
}
}  // extern "C"
