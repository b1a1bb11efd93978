// Pairing-family kernels: Pair, multi-pairing, PairingCheck, Miller loop, final exponentiation, G2 line tables.
// One batch element per thread.  AoS operands are read with 128-bit loads; every element is 64/128/384 B so a warp
// touches a contiguous 2-12 KB span (fully used sectors).
#include <stdlib.h>

#include "kcommon.cuh"
#include "pairing.cuh"

namespace bn254 {
namespace {
constexpr int kPairChunk = 4;  // pairs per shared-squaring pass inside one thread
#ifdef BN254_SMEM_SCRATCH
// per-thread Fp2 scratch of the tower routines (+ a tail when a timing-only probe squeezes the stride below 9 slots)
constexpr size_t kTowerSmem = (size_t)kBlock * kScratchStride + (kScratchSlots * 64 > kScratchStride ? kScratchSlots * 64 - kScratchStride + 64 : 0);
#else
constexpr size_t kTowerSmem = 0;
#endif
constexpr int kMpChunk = launch::kMpChunk;
static_assert(kLinesPerPoint == launch::kLinesPerPoint && sizeof(Fp2) * 3 == launch::kLineBytes, "launch.h out of date");

// Miller product of k pairs for one batch element, in passes of kPairChunk pairs.
// UNIFORM: the caller guarantees that every thread of the CTA is live and walks the same (k, chunk) schedule; the
// CTA then votes per pass whether any pair holds a point at infinity (the only data-dependent branch of the
// Miller loop) and runs the pass in lockstep when none does.
template <bool UNIFORM>
__device__ void miller_product(Fp12& f, const void* P, const void* Q, size_t first, int k) {
  G1Aff p[kPairChunk];
  G2Aff q[kPairChunk];
  G2Proj T[kPairChunk];
  bool have = false;
  for (int base = 0; base < k; base += kPairChunk) {
    int c = min(kPairChunk, k - base);
    bool finite = true;
    for (int j = 0; j < c; j++) {
      load_struct(p[j], P, first + base + j); load_struct(q[j], Q, first + base + j);
      finite = finite && !g1_is_inf(p[j]) && !g2_is_inf(q[j]);
    }
    if (UNIFORM) cta_lockstep_set(__syncthreads_and(finite) != 0);
    Fp12 g;
    Fp12& dst = have ? g : f;
    if (c == kPairChunk) miller_loop_t<kPairChunk>(dst, p, q, T, c);  // full passes: compile-time pair count
    else miller_loop_t<0>(dst, p, q, T, c);
    if (have) fp12_mul(f, f, g);
    have = true;
  }
}

// Coalesced CTA-wide staging: the kBlock operands of a CTA are contiguous in the caller's AoS arrays, so the CTA
// copies them with unit-stride 128-bit accesses (every warp instruction touches one contiguous 512-byte span)
// through the dynamic shared memory that later serves as the tower scratch, and each thread then picks its own
// element out of shared memory.  Same for the 384-byte results on the way out.
template <typename T>
__device__ __forceinline__ void cta_load(T& dst, const void* base, size_t first, size_t n_left, uint4* stage) {
  constexpr int Q4 = (int)(sizeof(T) / 16);
  const uint4* src = reinterpret_cast<const uint4*>(static_cast<const char*>(base) + first * sizeof(T));
  int total = (int)min((size_t)kBlock, n_left) * Q4;
  for (int w = threadIdx.x; w < total; w += kBlock) stage[w] = __ldg(src + w);
  __syncthreads();
  uint4* d = reinterpret_cast<uint4*>(&dst);
  if ((size_t)threadIdx.x < n_left) {
#pragma unroll
    for (int c = 0; c < Q4; c++) d[c] = stage[threadIdx.x * Q4 + c];
  }
  __syncthreads();
}
template <typename T>
__device__ __forceinline__ void cta_store(void* base, size_t first, size_t n_left, const T& src, uint4* stage) {
  constexpr int Q4 = (int)(sizeof(T) / 16);
  __syncthreads();  // the scratch is free again: every thread is past its last tower routine
  const uint4* sv = reinterpret_cast<const uint4*>(&src);
  if ((size_t)threadIdx.x < n_left) {
#pragma unroll
    for (int c = 0; c < Q4; c++) stage[threadIdx.x * Q4 + c] = sv[c];
  }
  __syncthreads();
  uint4* dst = reinterpret_cast<uint4*>(static_cast<char*>(base) + first * sizeof(T));
  int total = (int)min((size_t)kBlock, n_left) * Q4;
  for (int w = threadIdx.x; w < total; w += kBlock) dst[w] = stage[w];
}
// Start stagger (cycles, CTA-uniform): see launch::pair.
__device__ __forceinline__ void cta_stagger(unsigned stagger, unsigned mode) {
  if (stagger == 0) return;
  unsigned slot = mode == 0 ? blockIdx.x % 3u : ((blockIdx.x * 2654435761u) >> 16) % 16u;
  long long wait = (long long)slot * stagger, t0 = clock64();
  while (clock64() - t0 < wait) { }
}
__global__ void __launch_bounds__(kBlock, BN254_MIN_BLOCKS) k_pair(const void* P, const void* Q, size_t n, void* out, unsigned stagger, unsigned stagger_mode) {
  cta_stagger(stagger, stagger_mode);
  size_t first = (size_t)blockIdx.x * blockDim.x;
  size_t i = first + threadIdx.x;
  G1Aff p; G2Aff q; G2Proj T;
  Fp12 f;
#ifdef BN254_SMEM_SCRATCH
  cta_load(p, P, first, n - first, bn_dyn_smem);
  cta_load(q, Q, first, n - first, bn_dyn_smem);
  // lockstep (BN254_CTA_LOCKSTEP) only when all threads of the CTA take the same path: full CTA, no infinity
  bool plain = i < n && !g1_is_inf(p) && !g2_is_inf(q);
  cta_lockstep_set(__syncthreads_and(plain) != 0);
  if (i < n) {
    miller_loop(f, &p, &q, &T, 1);
    final_exp(f, f);
  }
  cta_store(out, first, n - first, f, bn_dyn_smem);
#else
  if (i >= n) return;
  load_struct(p, P, i); load_struct(q, Q, i);
  miller_loop(f, &p, &q, &T, 1);
  final_exp(f, f);
  store_struct(out, i, f);
#endif
}
// small products with a compile-time pair count (BLS verify: KC = 2): the pair loop unrolls
template <int MODE, int KC>
__global__ void __launch_bounds__(kBlock, BN254_MIN_BLOCKS) k_multi_pair_c(const void* P, const void* Q, size_t n, void* out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  G1Aff p[KC]; G2Aff q[KC]; G2Proj T[KC];
  bool plain = i < n;
  if (plain) {
#pragma unroll
    for (int j = 0; j < KC; j++) {
      load_struct(p[j], P, i * KC + j); load_struct(q[j], Q, i * KC + j);
      plain = plain && !g1_is_inf(p[j]) && !g2_is_inf(q[j]);
    }
  }
  cta_lockstep_set(__syncthreads_and(plain) != 0);  // lockstep: full CTA without points at infinity
  if (i >= n) return;
  Fp12 f;
  miller_loop_t<KC>(f, p, q, T, KC);
  if (MODE >= 1) final_exp(f, f);
  if (MODE == 2) static_cast<uint8_t*>(out)[i] = fp12_is_one(f) ? 1 : 0;
  else store_struct(out, i, f);
}
// PairingCheck of e(P0, Q0[i]) e(P1, Q1[i]) with the two G1 points shared by the whole batch: the shape of BLS
// verification (signature/bls01_signature/bls_signature.go:71-89: P0 = pk, P1 = -g1, Q0 = H(m_i), Q1 = sigma_i).
// Saves a third of the host->device bytes and the host-side replication of (pk, -g1) per message.
__global__ void __launch_bounds__(kBlock, BN254_MIN_BLOCKS) k_check2_fixed_g1(const void* P01, const void* Q0, const void* Q1, size_t n, uint8_t* ok) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  G1Aff p[2]; G2Aff q[2]; G2Proj T[2];
  bool plain = i < n;
  if (plain) {
    load_struct(p[0], P01, 0); load_struct(p[1], P01, 1);
    load_struct(q[0], Q0, i); load_struct(q[1], Q1, i);
    plain = !g1_is_inf(p[0]) && !g1_is_inf(p[1]) && !g2_is_inf(q[0]) && !g2_is_inf(q[1]);
  }
  cta_lockstep_set(__syncthreads_and(plain) != 0);
  if (i >= n) return;
  Fp12 f;
  miller_loop_t<2>(f, p, q, T, 2);
  final_exp(f, f);
  ok[i] = fp12_is_one(f) ? 1 : 0;
}
// mode 0: Miller product only; 1: + final exponentiation; 2: pairing check (writes one byte)
template <int MODE>
__global__ void __launch_bounds__(kBlock, BN254_MIN_BLOCKS) k_multi_pair(const void* P, const void* Q, size_t n, int k, void* out) {
  cta_lockstep_set(false);
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fp12 f;
  if (cta_is_full(n)) { miller_product<true>(f, P, Q, i * (size_t)k, k); cta_lockstep_set(true); }
  else miller_product<false>(f, P, Q, i * (size_t)k, k);
  if (MODE >= 1) final_exp(f, f);
  if (MODE == 2) static_cast<uint8_t*>(out)[i] = fp12_is_one(f) ? 1 : 0;
  else store_struct(out, i, f);
}
__global__ void __launch_bounds__(kBlock, BN254_MIN_BLOCKS) k_mp_partial(const void* P, const void* Q, size_t n, int k, int nchunks, void* partial) {
  // grid: x = blocks of kBlock products, y = pair group: every thread of a CTA walks the same number of pairs
  cta_lockstep_set(false);
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int ci = blockIdx.y;
  int first = ci * kMpChunk, cnt = min(kMpChunk, k - first);
  Fp12 f;
  if (cta_is_full(n)) miller_product<true>(f, P, Q, i * (size_t)k + first, cnt);
  else miller_product<false>(f, P, Q, i * (size_t)k + first, cnt);
  store_struct(partial, i * (size_t)nchunks + ci, f);
}
template <int MODE>
__global__ void __launch_bounds__(kBlock, BN254_MIN_BLOCKS) k_mp_combine(const void* partial, size_t n, int nchunks, void* out) {
  cta_lockstep_set(cta_is_full(n));
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fp12 f, g;
  load_struct(f, partial, i * (size_t)nchunks);
  for (int c = 1; c < nchunks; c++) { load_struct(g, partial, i * (size_t)nchunks + c); fp12_mul(f, f, g); }
  if (MODE >= 1) final_exp(f, f);
  if (MODE == 2) static_cast<uint8_t*>(out)[i] = fp12_is_one(f) ? 1 : 0;
  else store_struct(out, i, f);
}
// ---- precomputed G2 lines (fixed G2 points: user keys / public parameters) ---------------------------------
__global__ void __launch_bounds__(kBlock, BN254_MIN_BLOCKS) k_g2_lines(const void* Q, size_t m, Fp2* table, uint8_t* qskip) {
  cta_lockstep_set(false);
  size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  G2Aff q; load_struct(q, Q, j);
  bool inf = g2_is_inf(q);
  qskip[j] = inf ? 1 : 0;
  if (!inf) g2_precompute_lines(q, table + j * (size_t)kLinesPerPoint * 3);
}
#ifndef BN254_LINES_TMA
#define BN254_LINES_TMA 1
#endif
#if BN254_LINES_TMA && defined(BN254_CTA_LOCKSTEP)
// Line coefficients through shared memory (north_star: the Miller loop "reads precomputed G2 line coefficients for the
// schemes' fixed public-parameter G2 points from shared memory").  Every thread of a CTA walks the SAME table lines in
// the same order, so ONE thread fetches each couple of lines (2 x 192 B, contiguous per line) with the bulk-copy engine
// (cp.async.bulk = 1-D TMA) into a two-deep ring in shared memory, one application ahead of its use, and signals an
// mbarrier with the byte count; the 128 threads then read the coefficients as shared-memory broadcasts.  Fetch q of a
// CTA lands in buffer q & 1 and completes phase (q >> 1) & 1 of that buffer's mbarrier.  The ring holds 768 B per CTA:
// 3 CTAs x (74 KB scratch + 1 KB reserved + ring) still fit the 228 KB of an SM.  Only CTAs running in lockstep (full,
// no pair skipped -- the CTA barrier is what makes a shared ring safe) take this path; ragged CTAs read the table directly.
__shared__ alignas(128) Fp2 bn_line_ring[2][6];
__shared__ alignas(8) unsigned long long bn_line_bar[2];
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void line_ring_init() {  // one thread; the caller separates it from the first fetch by a CTA barrier
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&bn_line_bar[0])) : "memory");
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&bn_line_bar[1])) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// fetch q = schedule position q / F, couple q % F of the chunk's `cnt` table points (F couples per position, the last one
// a single line when cnt is odd)
__device__ __forceinline__ void line_ring_fetch(const Fp2* table, int first, int cnt, int q) {
  const int F = (cnt + 1) >> 1, s = q / F, ja = 2 * (q - s * F);
  const int nl = ja + 1 < cnt ? 2 : 1;
  const uint32_t bar = smem_addr(&bn_line_bar[q & 1]);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(nl * (int)launch::kLineBytes) : "memory");
  for (int l = 0; l < nl; l++) {
    const Fp2* src = table + ((size_t)(first + ja + l) * kLinesPerPoint + s) * 3;
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(&bn_line_ring[q & 1][3 * l])), "l"(src), "r"((int)launch::kLineBytes), "r"(bar) : "memory");
  }
}
__device__ __forceinline__ void line_ring_wait(int q) {
  const uint32_t bar = smem_addr(&bn_line_bar[q & 1]), parity = (uint32_t)(q >> 1) & 1u;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LINE_RING_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra LINE_RING_DONE;\n"
      "bra LINE_RING_WAIT;\n"
      "LINE_RING_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
// One schedule position of the line-table Miller loop out of the ring (lockstep CTAs: every pair live).
__device__ void miller_lines_step_ring(Fp12& f, const G1Aff* p, const Fp2* table, int first, int cnt, int& q, int nfetch, Fp2* sc) {
  for (int j = 0; j < cnt; j += 2, q++) {
    __syncthreads();  // every thread is done with the buffer the next fetch overwrites (it was read one application ago)
    if (threadIdx.x == 0 && q + 1 < nfetch) line_ring_fetch(table, first, cnt, q + 1);
    line_ring_wait(q);
    const Fp2* L = bn_line_ring[q & 1];
    if (j + 1 < cnt) apply_line_pair_mem(f, p[j], L, p[j + 1], L + 3, sc);
    else apply_line_mem(f, p[j], L, sc);
  }
}
#define BN254_LINES_RING 1
#else
#define BN254_LINES_RING 0
#endif
// Partial Miller products from line tables.  Grid: x = blocks of kBlock items, y = groups of kMpChunk pairs.
// Every thread of a CTA walks the SAME pairs: lockstep CTAs read each line out of the shared-memory ring above (one
// bulk copy per CTA), the others with warp-uniform (broadcast) loads of 192 B; P[i][j] is the only per-thread operand.
// out: partial[i * nchunks + chunk].
__global__ void __launch_bounds__(kBlock, BN254_MIN_BLOCKS) k_miller_lines(const void* P, const Fp2* __restrict__ table, const uint8_t* __restrict__ qskip,
                                                                          size_t n, int m, int nchunks, void* partial) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  int ci = blockIdx.y;
  int first = ci * launch::kLinesChunk, cnt = min(launch::kLinesChunk, m - first);
  G1Aff p[launch::kLinesChunk];
  unsigned skip = 0;
  if (i < n) {
    for (int j = 0; j < cnt; j++) {
      load_struct(p[j], P, i * (size_t)m + first + j);
      if (g1_is_inf(p[j]) || qskip[first + j]) skip |= 1u << j;
    }
  }
  cta_lockstep_set(__syncthreads_and(i < n && skip == 0) != 0);  // lockstep: full CTA, no pair skipped
  if (i >= n) return;
  Fp12 f;
  fp12_set_one(f);
  BN_SCRATCH_DECL
  int s = 0;
#if BN254_LINES_RING
  const bool ring = cta_lockstep_on();  // CTA-uniform
  const int nfetch = kLinesPerPoint * ((cnt + 1) >> 1);
  int q = 0;
  if (ring) {
    if (threadIdx.x == 0) line_ring_init();
    __syncthreads();
    if (threadIdx.x == 0) line_ring_fetch(table, first, cnt, 0);
  }
#endif
  for (int it = ATE_NAF_LEN - 2; it >= -2; it--) {
    // it >= 0: tangent (+ chord if the digit is non-zero); it == -1, -2: the two Frobenius lines
    if (it >= 0 && it != ATE_NAF_LEN - 2) fp12_sqr(f, f);
    int reps = (it >= 0 && ATE_NAF[it]) ? 2 : 1;
    for (int r = 0; r < reps; r++, s++) {
#if BN254_LINES_RING
      if (ring) { miller_lines_step_ring(f, p, table, first, cnt, q, nfetch, sc_); continue; }
#endif
      miller_lines_step(f, p, table, first, cnt, skip, s, sc_);
    }
  }
  store_struct(partial, i * (size_t)nchunks + ci, f);
}
__global__ void __launch_bounds__(kBlock, BN254_MIN_BLOCKS) k_final_exp(const void* in, size_t n, void* out) {
  cta_lockstep_set(cta_is_full(n));
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fp12 f; load_struct(f, in, i);
  final_exp(f, f);
  store_struct(out, i, f);
}
// expands the fixed-G1 check into the (P, Q) pair arrays the per-pair lane-group Miller kernel reads
__global__ void k_pack_check2(const void* P01, const void* Q0, const void* Q1, size_t n, void* P, void* Q) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  G1Aff p0, p1; G2Aff q0, q1;
  load_struct(p0, P01, 0); load_struct(p1, P01, 1); load_struct(q0, Q0, i); load_struct(q1, Q1, i);
  store_struct(P, 2 * i, p0); store_struct(P, 2 * i + 1, p1);
  store_struct(Q, 2 * i, q0); store_struct(Q, 2 * i + 1, q1);
}
// ok[i] = (x[i] == 1): the comparison half of PairingCheck when the final exponentiation ran in another kernel
__global__ void k_gt_is_one(const void* x, size_t n, uint8_t* ok) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fp12 f; load_struct(f, x, i);
  ok[i] = fp12_is_one(f) ? 1 : 0;
}

template <int MODE>
void launch_multi_pair(const void* P, const void* Q, size_t n, int k, void* out, cudaStream_t s) {
  if (k == 1) BN_LAUNCH, k_multi_pair_c<MODE, 1><<<grid_for(n), kBlock, kTowerSmem, s>>>(P, Q, n, out);
  else if (k == 2) BN_LAUNCH, k_multi_pair_c<MODE, 2><<<grid_for(n), kBlock, kTowerSmem, s>>>(P, Q, n, out);
  else if (k == 3) BN_LAUNCH, k_multi_pair_c<MODE, 3><<<grid_for(n), kBlock, kTowerSmem, s>>>(P, Q, n, out);
  else BN_LAUNCH, k_multi_pair<MODE><<<grid_for(n), kBlock, kTowerSmem, s>>>(P, Q, n, k, out);
}

}  // namespace

namespace launch {

cudaError_t pairing_init() {
#ifdef BN254_SMEM_SCRATCH
  const void* kernels[] = {(const void*)k_pair, (const void*)k_g2_lines, (const void*)k_miller_lines,
                           (const void*)k_multi_pair_c<0, 1>, (const void*)k_multi_pair_c<1, 1>, (const void*)k_multi_pair_c<2, 1>,
                           (const void*)k_multi_pair_c<0, 2>, (const void*)k_multi_pair_c<1, 2>, (const void*)k_multi_pair_c<2, 2>,
                           (const void*)k_multi_pair_c<0, 3>, (const void*)k_multi_pair_c<1, 3>, (const void*)k_multi_pair_c<2, 3>,
                           (const void*)k_multi_pair<0>, (const void*)k_multi_pair<1>, (const void*)k_multi_pair<2>,
                           (const void*)k_mp_partial, (const void*)k_mp_combine<0>, (const void*)k_mp_combine<1>, (const void*)k_mp_combine<2>,
                           (const void*)k_final_exp, (const void*)k_check2_fixed_g1};
  for (const void* k : kernels) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTowerSmem);
    if (e != cudaSuccess) return e;
  }
#endif
  return cudaSuccess;
}
int pairing_wave_threads(int sms) { return sms * BN254_MIN_BLOCKS * kBlock; }

void pair(const void* P, const void* Q, size_t n, void* out, cudaStream_t s) {
  static const unsigned stagger = getenv("BN254_STAGGER") ? (unsigned)atoi(getenv("BN254_STAGGER")) : 0u;
  static const unsigned mode = getenv("BN254_STAGGER_MODE") ? (unsigned)atoi(getenv("BN254_STAGGER_MODE")) : 0u;
  BN_LAUNCH, k_pair<<<grid_for(n), kBlock, kTowerSmem, s>>>(P, Q, n, out, stagger, mode);
}
void multi_pair(int mode, const void* P, const void* Q, size_t n, int k, void* out, cudaStream_t s) {
  if (mode == 0) launch_multi_pair<0>(P, Q, n, k, out, s);
  else if (mode == 1) launch_multi_pair<1>(P, Q, n, k, out, s);
  else launch_multi_pair<2>(P, Q, n, k, out, s);
}
void mp_partial(const void* P, const void* Q, size_t n, int k, int nchunks, void* partial, cudaStream_t s) {
  BN_LAUNCH, k_mp_partial<<<dim3(grid_for(n), (unsigned)nchunks), kBlock, kTowerSmem, s>>>(P, Q, n, k, nchunks, partial);
}
void mp_combine(int mode, const void* partial, size_t n, int nchunks, void* out, cudaStream_t s) {
  if (mode == 0) BN_LAUNCH, k_mp_combine<0><<<grid_for(n), kBlock, kTowerSmem, s>>>(partial, n, nchunks, out);
  else if (mode == 1) BN_LAUNCH, k_mp_combine<1><<<grid_for(n), kBlock, kTowerSmem, s>>>(partial, n, nchunks, out);
  else BN_LAUNCH, k_mp_combine<2><<<grid_for(n), kBlock, kTowerSmem, s>>>(partial, n, nchunks, out);
}
void g2_lines(const void* Q, size_t m, void* table, uint8_t* qskip, cudaStream_t s) {
  BN_LAUNCH, k_g2_lines<<<grid_for(m), kBlock, kTowerSmem, s>>>(Q, m, static_cast<Fp2*>(table), qskip);
}
void miller_lines(const void* P, const void* table, const uint8_t* qskip, size_t n, int m, int nchunks, void* partial, cudaStream_t s) {
  BN_LAUNCH, k_miller_lines<<<dim3(grid_for(n), (unsigned)nchunks), kBlock, kTowerSmem, s>>>(P, static_cast<const Fp2*>(table), qskip, n, m, nchunks, partial);
}
void final_exp(const void* in, size_t n, void* out, cudaStream_t s) { BN_LAUNCH, k_final_exp<<<grid_for(n), kBlock, kTowerSmem, s>>>(in, n, out); }
void check2_fixed_g1(const void* P01, const void* Q0, const void* Q1, size_t n, uint8_t* ok, cudaStream_t s) {
  BN_LAUNCH, k_check2_fixed_g1<<<grid_for(n), kBlock, kTowerSmem, s>>>(P01, Q0, Q1, n, ok);
}
void pack_check2(const void* P01, const void* Q0, const void* Q1, size_t n, void* P, void* Q, cudaStream_t s) {
  BN_LAUNCH, k_pack_check2<<<grid_for(n), kBlock, 0, s>>>(P01, Q0, Q1, n, P, Q);
}
void gt_is_one(const void* x, size_t n, uint8_t* ok, cudaStream_t s) { BN_LAUNCH, k_gt_is_one<<<grid_for(n), kBlock, 0, s>>>(x, n, ok); }

}  // namespace launch
}  // namespace bn254
