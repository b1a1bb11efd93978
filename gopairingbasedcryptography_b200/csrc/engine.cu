// libbn254_b200.so -- kernels and host runtime behind include/bn254_b200.h.
// sm_100a only; no CPU fallback: every entry point needs a live CUDA device.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <algorithm>

#include "../../include/bn254_b200.h"
#include "curve.cuh"
#include "hash_to_curve.cuh"
#include "vm.cuh"

using namespace bn254;

// =============================================================================================
// Kernels: one batch element per thread.  AoS operands are read with 128-bit loads; every element
// is 64/128/384 B so a warp touches a contiguous 2-12 KB span (fully used sectors).
// =============================================================================================
namespace {

#ifndef BN254_BLOCK
#define BN254_BLOCK 128
#endif
constexpr int kBlock = BN254_BLOCK;
#ifndef BN254_MIN_BLOCKS
#define BN254_MIN_BLOCKS 1
#endif
constexpr int kPairChunk = 4;
#ifdef BN254_SMEM_SCRATCH
constexpr size_t kTowerSmem = (size_t)kBlock * kScratchStride;  // per-thread Fp2 scratch of the tower routines
#else
constexpr size_t kTowerSmem = 0;
#endif  // pairs per shared-squaring pass inside one thread

template <typename T>
__device__ __forceinline__ void load_struct(T& dst, const void* base, size_t idx) {
  static_assert(sizeof(T) % 16 == 0, "16-byte multiple");
  const uint4* src = reinterpret_cast<const uint4*>(static_cast<const char*>(base) + idx * sizeof(T));
  uint4* d = reinterpret_cast<uint4*>(&dst);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(T) / 16); i++) d[i] = __ldg(src + i);
}
template <typename T>
__device__ __forceinline__ void store_struct(void* base, size_t idx, const T& src) {
  uint4* dst = reinterpret_cast<uint4*>(static_cast<char*>(base) + idx * sizeof(T));
  const uint4* s = reinterpret_cast<const uint4*>(&src);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(T) / 16); i++) dst[i] = s[i];
}

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
__device__ __forceinline__ bool cta_is_full(size_t n) { return ((size_t)blockIdx.x + 1) * blockDim.x <= n; }

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
__global__ void __launch_bounds__(kBlock, BN254_MIN_BLOCKS) k_pair(const void* P, const void* Q, size_t n, void* out) {
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
// Large products (BSW07-style decryption, k ~ 200 pairs): the k pairs of one product are split into groups of
// kMpChunk pairs, one thread per group (n * ceil(k/kMpChunk) threads), then k_mp_combine multiplies the
// partial Miller values of a product and finishes with ONE final exponentiation / check.
constexpr int kMpChunk = 8;
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
// Partial Miller products from line tables.  Grid: x = blocks of kBlock items, y = groups of kMpChunk pairs.
// Every thread of a CTA walks the SAME pairs, so each line read is one warp-uniform (broadcast) load of 192 B
// served by L1; P[i][j] is the only per-thread operand.  out: partial[i * nchunks + chunk].
__global__ void __launch_bounds__(kBlock, BN254_MIN_BLOCKS) k_miller_lines(const void* P, const Fp2* __restrict__ table, const uint8_t* __restrict__ qskip,
                                                                          size_t n, int m, int nchunks, void* partial) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  int ci = blockIdx.y;
  int first = ci * kMpChunk, cnt = min(kMpChunk, m - first);
  G1Aff p[kMpChunk];
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
  for (int it = ATE_NAF_LEN - 2; it >= -2; it--) {
    // it >= 0: tangent (+ chord if the digit is non-zero); it == -1, -2: the two Frobenius lines
    if (it >= 0 && it != ATE_NAF_LEN - 2) fp12_sqr(f, f);
    int reps = (it >= 0 && ATE_NAF[it]) ? 2 : 1;
    for (int r = 0; r < reps; r++, s++) {
      for (int j = 0; j < cnt; j++) {
        if ((skip >> j) & 1u) continue;
        const Fp2* L = table + ((size_t)(first + j) * kLinesPerPoint + s) * 3;
        apply_line_mem(f, p[j], L, sc_);
      }
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
template <typename J, typename A>
__global__ void __launch_bounds__(kBlock, BN254_MIN_BLOCKS) k_scalar_mul(const void* base, size_t base_stride, const void* scalars, size_t n, void* out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  A b;
  bool plain = i < n;
  if (plain) { load_struct(b, base, i * base_stride); plain = !aff_is_inf(b); }
  cta_lockstep_set(__syncthreads_and(plain) != 0);  // the ladder has a fixed trip count; infinity bases return early
  if (i >= n) return;
  uint32_t s[8];
  const uint4* sp = reinterpret_cast<const uint4*>(static_cast<const char*>(scalars) + i * 32);
  uint4 lo = __ldg(sp), hi = __ldg(sp + 1);
  s[0] = lo.x; s[1] = lo.y; s[2] = lo.z; s[3] = lo.w; s[4] = hi.x; s[5] = hi.y; s[6] = hi.z; s[7] = hi.w;
  A r;
  Fp beta = (sizeof(A) == sizeof(G1Aff)) ? GLV_BETA : GLV_BETA_G2;
  scalar_mul_glv<J, A>(r, b, s, beta);
  store_struct(out, i, r);
}
// fixed base: 32 windowed mixed additions from a precomputed affine table (L2-resident, 0.5-1 MB)
template <typename J, typename A>
__global__ void __launch_bounds__(kBlock, BN254_MIN_BLOCKS) k_fixed_mul(const A* table, const void* scalars, size_t n, void* out) {
  cta_lockstep_set(false);  // no barriers in this kernel; the flag is read by the shared field routines
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t s[8];
  const uint4* sp = reinterpret_cast<const uint4*>(static_cast<const char*>(scalars) + i * 32);
  uint4 lo = __ldg(sp), hi = __ldg(sp + 1);
  s[0] = lo.x; s[1] = lo.y; s[2] = lo.z; s[3] = lo.w; s[4] = hi.x; s[5] = hi.y; s[6] = hi.z; s[7] = hi.w;
  A r;
  scalar_mul_fixed<J, A>(r, table, s);
  store_struct(out, i, r);
}
template <typename J, typename A>
__global__ void __launch_bounds__(kBlock, BN254_MIN_BLOCKS) k_aff_add(const void* a, const void* b, size_t n, void* out) {
  cta_lockstep_set(false);  // no barriers in this kernel; the flag is read by the shared field routines
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  A x, y, r; load_struct(x, a, i); load_struct(y, b, i);
  aff_add<J, A>(r, x, y);
  store_struct(out, i, r);
}
// out[i] = U[0] + sum_{j < m, bit j of sel_i set} U[j+1]   (Waters hash: ibe/waters05_ibe/waters05_ibe.go:227-233).
// Bit j is bit (7 - j%8) of byte j/8 -- the MSB-first order of waters05_ibe.go:302-313.  The m+1 public
// points are staged in shared memory once per CTA; the sum runs in Jacobian form with ONE inversion at the end
// (the reference pays one inversion per Add).
template <typename J, typename A>
__global__ void __launch_bounds__(kBlock, BN254_MIN_BLOCKS) k_subset_sum(const A* U, int m, const uint8_t* sel, size_t n, void* out) {
  cta_lockstep_set(false);  // no barriers in this kernel; the flag is read by the shared field routines
  extern __shared__ uint4 su_raw[];
  A* su = reinterpret_cast<A*>(su_raw);
  {
    const uint4* src = reinterpret_cast<const uint4*>(U);
    int words = (m + 1) * (int)(sizeof(A) / 16);
    for (int w = threadIdx.x; w < words; w += blockDim.x) su_raw[w] = __ldg(src + w);
  }
  __syncthreads();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint8_t* bits = sel + i * (size_t)((m + 7) / 8);
  J acc;
  if (aff_is_inf(su[0])) { f_set_zero(acc.x); f_set_zero(acc.y); f_set_zero(acc.z); }
  else { acc.x = su[0].x; acc.y = su[0].y; f_set_one(acc.z); }
  for (int j = 0; j < m; j++) {
    if ((bits[j >> 3] >> (7 - (j & 7))) & 1) {
      A e = su[j + 1];
      if (!aff_is_inf(e)) jac_add_aff(acc, acc, e);
    }
  }
  A r;
  jac_to_aff(r, acc);
  store_struct(out, i, r);
}
// out[g] = sum of the `len` consecutive points of group g, processed as ceil(len/32)-way partial sums per pass
template <typename J, typename A>
__global__ void __launch_bounds__(kBlock, BN254_MIN_BLOCKS) k_segment_sum(const void* pts, size_t groups, int len, int chunk, void* out) {
  cta_lockstep_set(false);  // no barriers in this kernel; the flag is read by the shared field routines
  int nch = (len + chunk - 1) / chunk;
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= groups * (size_t)nch) return;
  size_t g = t / nch;
  int c = (int)(t % nch);
  int first = c * chunk, cnt = min(chunk, len - first);
  J acc; f_set_zero(acc.x); f_set_zero(acc.y); f_set_zero(acc.z);
  for (int j = 0; j < cnt; j++) {
    A e; load_struct(e, pts, g * (size_t)len + first + j);
    if (!aff_is_inf(e)) jac_add_aff(acc, acc, e);
  }
  A r;
  jac_to_aff(r, acc);
  store_struct(out, t, r);
}
constexpr size_t kGtCycloTable = 16;  // Fp12 entries of per-thread table space gt_cyclo_exp needs
template <int CYCLO>
__global__ void __launch_bounds__(kBlock, BN254_MIN_BLOCKS) k_gt_exp(const void* x, size_t x_stride, const void* k, size_t n, void* out, Fp12* tabmem = nullptr) {
  cta_lockstep_set(cta_is_full(n));  // gt_exp / gt_cyclo_exp have thread-uniform control flow
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fp12 b; load_struct(b, x, i * x_stride);
  uint32_t s[8];
  const uint4* sp = reinterpret_cast<const uint4*>(static_cast<const char*>(k) + i * 32);
  uint4 lo = __ldg(sp), hi = __ldg(sp + 1);
  s[0] = lo.x; s[1] = lo.y; s[2] = lo.z; s[3] = lo.w; s[4] = hi.x; s[5] = hi.y; s[6] = hi.z; s[7] = hi.w;
  Fp12 r;
  if (CYCLO) gt_cyclo_exp(r, b, s, tabmem + i * kGtCycloTable);
  else {
    Fp12 loc[4];  // table on the stack when the launch has no scratch (device-pointer entry point, table builds)
    gt_exp(r, b, s, tabmem ? tabmem + i * 4 : loc);
  }
  store_struct(out, i, r);
}
// fixed-base GT exponentiation: out = prod_w table[w][byte_w(k)] -- 32 Fp12 products, no squarings.  The table
// (32 x 255 x 384 B = 3.1 MB, L2-resident) is built once per base with k_gt_exp on the scalars d << 8w.
__global__ void __launch_bounds__(kBlock, BN254_MIN_BLOCKS) k_gt_fixed_exp(const Fp12* table, const void* k, size_t n, void* out) {
  cta_lockstep_set(cta_is_full(n));
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t s[8];
  const uint4* sp = reinterpret_cast<const uint4*>(static_cast<const char*>(k) + i * 32);
  uint4 lo = __ldg(sp), hi = __ldg(sp + 1);
  s[0] = lo.x; s[1] = lo.y; s[2] = lo.z; s[3] = lo.w; s[4] = hi.x; s[5] = hi.y; s[6] = hi.z; s[7] = hi.w;
  Fp12 acc, e;
  fp12_set_one(acc);
  for (int w = 0; w < kFixedWindows; w++) {  // uniform: one product per window, by 1 when the digit is 0
    int d = (int)((s[w >> 2] >> ((w & 3) * 8)) & 0xFFu);
    if (d) load_struct(e, table, (size_t)w * kFixedEntries + d - 1);
    else fp12_set_one(e);
    fp12_mul(acc, acc, e);
  }
  store_struct(out, i, acc);
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
// mode 0: a*b ; mode 1: a/b
template <int MODE>
__global__ void __launch_bounds__(kBlock, BN254_MIN_BLOCKS) k_gt_mul(const void* a, const void* b, size_t n, void* out) {
  cta_lockstep_set(false);
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fp12 x, y; load_struct(x, a, i); load_struct(y, b, i);
  if (MODE == 1) fp12_inv(y, y);
  fp12_mul(x, x, y);
  store_struct(out, i, x);
}
// hash-to-curve: one message per thread (SHA-256 expand_message_xmd, SVDW map x2, add, G2 cofactor clearing)
template <int G>
__global__ void __launch_bounds__(kBlock, BN254_MIN_BLOCKS) k_hash_to_curve(const uint8_t* msgs, const uint64_t* off, size_t n, const uint8_t* dst,
                                                                           uint32_t dst_len, void* out) {
  cta_lockstep_set(false);  // no barriers in this kernel; the flag is read by the shared field routines
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint8_t* m = msgs + off[i];
  size_t len = (size_t)(off[i + 1] - off[i]);
  if (G == 1) { G1Aff r; hash_to_g1(r, m, len, dst, dst_len); store_struct(out, i, r); }
  else { G2Aff r; hash_to_g2(r, m, len, dst, dst_len); store_struct(out, i, r); }
}
__global__ void __launch_bounds__(kBlock, BN254_MIN_BLOCKS) k_fp_mul(const void* a, const void* b, size_t n, void* out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fp x, y; load_struct(x, a, i); load_struct(y, b, i);
  x = fp_mul(x, y);
  store_struct(out, i, x);
}


// ---------------------------------------------------------------------------------------------
// Lane-group ("tower VM") kernels: K lanes per pairing, state in shared memory.  See vm.cuh.
// ---------------------------------------------------------------------------------------------
#ifndef BN254_VM_K
#define BN254_VM_K 3
#endif
#ifndef BN254_VM_WARPS
#define BN254_VM_WARPS 4
#endif
constexpr int kVmK = BN254_VM_K;
constexpr int kVmWarps = BN254_VM_WARPS;
constexpr int kVmGroups = 32 / kVmK;            // pairings per warp
constexpr int kVmNP = kVmGroups * kVmWarps;     // pairings per CTA
constexpr int kVmStride = kVmNP | 1;            // odd stride: sub-lanes of a group hit different bank quads
constexpr int kVmColdSlots = 64;                // cold slots reserved per pairing in the global scratch

#define VM_CAT_(a, b, c) a##b##c
#define VM_CAT(a, b, c) VM_CAT_(a, b, c)
#define VM_SYM(name, suffix) VM_CAT(vm::name##_K, BN254_VM_K, suffix)
#if BN254_VM_K == 1
#define VM_INC_PAIR "vm_prog_pair_k1.inc"
#define VM_INC_MILLER "vm_prog_miller_k1.inc"
#define VM_INC_FINALEXP "vm_prog_finalexp_k1.inc"
#elif BN254_VM_K == 2
#define VM_INC_PAIR "vm_prog_pair_k2.inc"
#define VM_INC_MILLER "vm_prog_miller_k2.inc"
#define VM_INC_FINALEXP "vm_prog_finalexp_k2.inc"
#elif BN254_VM_K == 3
#define VM_INC_PAIR "vm_prog_pair_k3.inc"
#define VM_INC_MILLER "vm_prog_miller_k3.inc"
#define VM_INC_FINALEXP "vm_prog_finalexp_k3.inc"
#elif BN254_VM_K == 4
#define VM_INC_PAIR "vm_prog_pair_k4.inc"
#define VM_INC_MILLER "vm_prog_miller_k4.inc"
#define VM_INC_FINALEXP "vm_prog_finalexp_k4.inc"
#elif BN254_VM_K == 6
#define VM_INC_PAIR "vm_prog_pair_k6.inc"
#define VM_INC_MILLER "vm_prog_miller_k6.inc"
#define VM_INC_FINALEXP "vm_prog_finalexp_k6.inc"
#else
#error "BN254_VM_K must be 1, 2, 3, 4 or 6"
#endif
__device__ const uint64_t kProgPair[] = {
#include VM_INC_PAIR
};
__device__ const uint64_t kProgMiller[] = {
#include VM_INC_MILLER
};
__device__ const uint64_t kProgFinalExp[] = {
#include VM_INC_FINALEXP
};

struct VmProgPair { static constexpr int rounds = VM_SYM(PAIR, _ROUNDS), nslots = VM_SYM(PAIR, _NSLOTS), nin = 3;
  __device__ static const uint64_t* prog() { return kProgPair; }
  __device__ static int in(int i) { return VM_SYM(PAIR, _IN)[i]; } __device__ static int out(int i) { return VM_SYM(PAIR, _OUT)[i]; } };
struct VmProgMiller { static constexpr int rounds = VM_SYM(MILLER, _ROUNDS), nslots = VM_SYM(MILLER, _NSLOTS), nin = 3;
  __device__ static const uint64_t* prog() { return kProgMiller; }
  __device__ static int in(int i) { return VM_SYM(MILLER, _IN)[i]; } __device__ static int out(int i) { return VM_SYM(MILLER, _OUT)[i]; } };
struct VmProgFinalExp { static constexpr int rounds = VM_SYM(FINALEXP, _ROUNDS), nslots = VM_SYM(FINALEXP, _NSLOTS), nin = 6;
  __device__ static const uint64_t* prog() { return kProgFinalExp; }
  __device__ static int in(int i) { return VM_SYM(FINALEXP, _IN)[i]; } __device__ static int out(int i) { return VM_SYM(FINALEXP, _OUT)[i]; } };

template <typename PROG> constexpr size_t vm_smem_bytes() { return (size_t)PROG::nslots * 4 * kVmStride * sizeof(uint4); }

// Persistent CTAs: each warp owns kVmGroups pairings at a time; warps never synchronise with each other.
// PROG::nin == 3: inputs are (P, Q.x, Q.y) from the G1/G2 arrays; PROG::nin == 6: the six Fp2 of a GT.
template <typename PROG>
__global__ void __launch_bounds__(32 * kVmWarps) k_vm(const void* in0, const void* in1, size_t n, void* out, uint4* cold) {
  extern __shared__ uint4 vm_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane / kVmK, j = lane % kVmK;
  const bool lane_ok = g < kVmGroups;
  const int pid = warp * kVmGroups + (lane_ok ? g : 0);
  vm::SlotFile f;
  f.hot = vm_smem; f.nslots = PROG::nslots; f.hot_stride = kVmStride; f.pid = pid;
  f.cold = cold; f.cold_stride = gridDim.x * kVmNP; f.gpid = blockIdx.x * kVmNP + pid;
  const unsigned gmask = lane_ok ? (((1u << kVmK) - 1u) << (g * kVmK)) : 0u;
  for (size_t base = (size_t)blockIdx.x * kVmNP; base < n; base += (size_t)gridDim.x * kVmNP) {
    const size_t idx = base + pid;
    const bool active = lane_ok && idx < n;
    // ---- prologue: operands -> slots; pairs containing the point at infinity are flagged ----
    unsigned nzP = 0, nzQ = 0;
    for (int i = 0; i < PROG::nin; i++) {
      bool mine = active && (i % kVmK) == j;
      uint32_t nz = 0;
      if (mine) {
        const char* src;
        if (PROG::nin == 3) src = (i == 0) ? static_cast<const char*>(in0) + idx * 64 : static_cast<const char*>(in1) + idx * 128 + (i - 1) * 64;
        else src = static_cast<const char*>(in0) + idx * 384 + i * 64;
        Fp2 v;
        uint4* d = reinterpret_cast<uint4*>(&v);
#pragma unroll
        for (int c = 0; c < 4; c++) { d[c] = __ldg(reinterpret_cast<const uint4*>(src) + c); nz |= d[c].x | d[c].y | d[c].z | d[c].w; }
        vm::st_slot(f, PROG::in(i), v);
      }
      unsigned b = __ballot_sync(0xffffffffu, nz != 0);
      if (i == 0) nzP = b & gmask; else nzQ |= b & gmask;
    }
    const bool skip = (PROG::nin == 3) && (nzP == 0 || nzQ == 0);
    __syncwarp();
    vm::run<kVmK>(f, PROG::prog(), PROG::rounds, j, active);
    // ---- epilogue ----
    if (active) {
      for (int i = j; i < 6; i += kVmK) {
        Fp2 v;
        if (skip) { v = fp2_zero(); if (i == 0) v.a0 = fp_one(); }
        else v = vm::ld_slot(f, PROG::out(i));
        uint4* dst = reinterpret_cast<uint4*>(static_cast<char*>(out) + idx * 384 + i * 64);
        const uint4* sv = reinterpret_cast<const uint4*>(&v);
#pragma unroll
        for (int c = 0; c < 4; c++) dst[c] = sv[c];
      }
    }
    __syncwarp();
  }
}

inline unsigned grid_for(size_t n) { return (unsigned)((n + kBlock - 1) / kBlock); }

}  // namespace

// =============================================================================================
// Host runtime
// =============================================================================================
struct Slot {
  cudaStream_t stream = nullptr;
  char* h = nullptr;  // pinned staging
  char* d = nullptr;  // device staging
  uint4* vm_cold = nullptr;  // cold slot scratch of the lane-group kernels launched on this stream
  void* mp_scratch = nullptr;  // partial Miller values of split multi-pairings launched on this stream
  size_t mp_scratch_bytes = 0;
  // pending output copy-back
  void* user_out = nullptr;
  size_t out_off = 0, out_bytes = 0;
  bool busy = false;
};

struct bn254_ctx {
  int device = 0;
  std::mutex mu;
  std::string err;
  Slot slot[2];
  size_t slot_bytes = 0;
  uint64_t launches = 0;
  // lane-group (tower VM) kernels
  // BN254_IMPL: "vm" = lane-group (tower VM) kernels always, "thread" = one-thread-per-pairing kernels always,
  // unset = thread kernels, except that launches of at most kVmAutoMax elements take the lane-group kernels: they
  // finish a small batch in 5.8 ms instead of 10.6 ms (profiles/r1/latency_vs_batch.jsonl; crossover ~20k elements)
  int vm_mode = 0;             // 0 auto, 1 always, 2 never
  int sms = 0;
  int vm_blocks_per_sm[3] = {0, 0, 0};  // pair, miller, finalexp
  uint4* vm_cold_dev = nullptr;  // scratch for the *_dev entry points (launches are serialised by vm_dev_done)
  // fixed-base window tables (one cached base per group), built on first use with the GLV kernel
  Slot dev_slot;  // only mp_scratch is used: scratch of the *_dev multi-pairing launches
  void* fixed_table[2] = {nullptr, nullptr};
  unsigned char fixed_base[2][BN254_G2_BYTES] = {};
  bool fixed_valid[2] = {false, false};
  Fp12* gt_table = nullptr;  // fixed-base GT table of the last base used with >= kFixedMin exponents
  unsigned char gt_base[BN254_GT_BYTES] = {};
  bool gt_valid = false;
  cudaEvent_t vm_dev_done = nullptr;
};

namespace {

constexpr size_t kSlotBytes = 96u << 20;       // per-slot staging (pinned + device)
constexpr size_t kMaxChunkItems = 1u << 17;   // keeps two chunks in flight for copy/compute overlap

int fail(bn254_ctx* c, int code, const char* what, cudaError_t e = cudaSuccess) {
  if (c) {
    c->err = what;
    if (e != cudaSuccess) { c->err += ": "; c->err += cudaGetErrorString(e); }
  }
  return code;
}
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(ctx, e_ == cudaErrorMemoryAllocation ? BN254_ERR_OOM : BN254_ERR_CUDA, #call, e_); } while (0)

struct Operand { const void* ptr; size_t item_bytes; bool broadcast; };
// Chunked, double-buffered host-buffer driver: up to three input operands (each per-item or broadcast), one output.
// launch(d_in[3], count, d_out, stream, cold)
template <typename L>
int run_host_n(bn254_ctx* ctx, const Operand* in, int nin, void* out, size_t out_item, size_t n, L launch) {
  if (!ctx) return BN254_ERR_BAD_ARG;
  if (n == 0) return BN254_OK;
  if (!out) return fail(ctx, BN254_ERR_BAD_ARG, "null pointer");
  size_t fixed = 1024, per = out_item;
  for (int k = 0; k < nin; k++) {
    if (in[k].item_bytes && !in[k].ptr) return fail(ctx, BN254_ERR_BAD_ARG, "null pointer");
    if (in[k].broadcast) fixed += in[k].item_bytes + 256; else per += in[k].item_bytes;
  }
  std::lock_guard<std::mutex> lk(ctx->mu);
  CU(cudaSetDevice(ctx->device));
  size_t chunk = std::min<size_t>({n, kMaxChunkItems, (ctx->slot_bytes - fixed) / per});
  if (chunk == 0) return fail(ctx, BN254_ERR_BAD_ARG, "element too large for staging");
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
  auto finish = [&](Slot& s) -> int {
    if (!s.busy) return BN254_OK;
    CU(cudaStreamSynchronize(s.stream));
    if (s.out_bytes) memcpy(s.user_out, s.h + s.out_off, s.out_bytes);
    s.busy = false;
    return BN254_OK;
  };
  size_t done = 0;
  int ci = 0;
  while (done < n) {
    size_t c = std::min(chunk, n - done);
    Slot& s = ctx->slot[ci & 1];
    int rc = finish(s);
    if (rc) return rc;
    size_t off = 0;
    const void* d_in[3] = {nullptr, nullptr, nullptr};
    for (int k = 0; k < nin; k++) {
      d_in[k] = s.d + off;
      if (!in[k].item_bytes) continue;
      size_t l = in[k].broadcast ? in[k].item_bytes : in[k].item_bytes * c;
      const char* src = static_cast<const char*>(in[k].ptr) + (in[k].broadcast ? 0 : done * in[k].item_bytes);
      if (in_pinned[k]) CU(cudaMemcpyAsync(s.d + off, src, l, cudaMemcpyHostToDevice, s.stream));
      else { memcpy(s.h + off, src, l); CU(cudaMemcpyAsync(s.d + off, s.h + off, l, cudaMemcpyHostToDevice, s.stream)); }
      off = (off + l + 255) & ~size_t(255);
    }
    size_t oo = off, lo = out_item * c;
    launch(d_in, c, s.d + oo, s.stream, s.vm_cold);
    ctx->launches++;
    CU(cudaGetLastError());
    s.user_out = static_cast<char*>(out) + done * out_item;
    if (out_pinned) { CU(cudaMemcpyAsync(s.user_out, s.d + oo, lo, cudaMemcpyDeviceToHost, s.stream)); s.out_bytes = 0; }
    else { CU(cudaMemcpyAsync(s.h + oo, s.d + oo, lo, cudaMemcpyDeviceToHost, s.stream)); s.out_bytes = lo; }
    s.out_off = oo; s.busy = true;
    done += c; ci++;
  }
  for (int i = 0; i < 2; i++) { int rc = finish(ctx->slot[(ci + i) & 1]); if (rc) return rc; }
  return BN254_OK;
}
// two-operand form: launch(d_in0, d_in1, count, d_out, stream, cold)
template <typename L>
int run_host(bn254_ctx* ctx, Operand in0, Operand in1, void* out, size_t out_item, size_t n, L launch) {
  if (n && !in0.ptr) return fail(ctx, BN254_ERR_BAD_ARG, "null pointer");
  Operand in[2] = {in0, in1};
  return run_host_n(ctx, in, 2, out, out_item, n,
                    [&](const void* const* d, size_t c, void* o, cudaStream_t s, uint4* cold) { launch(d[0], d[1], c, o, s, cold); });
}

template <typename L>
int run_dev(bn254_ctx* ctx, size_t n, L launch) {
  if (!ctx) return BN254_ERR_BAD_ARG;
  if (n == 0) return BN254_OK;
  std::lock_guard<std::mutex> lk(ctx->mu);
  CU(cudaSetDevice(ctx->device));
  launch();
  ctx->launches++;
  CU(cudaGetLastError());
  return BN254_OK;
}


template <typename PROG> constexpr int vm_prog_index();
template <> constexpr int vm_prog_index<VmProgPair>() { return 0; }
template <> constexpr int vm_prog_index<VmProgMiller>() { return 1; }
template <> constexpr int vm_prog_index<VmProgFinalExp>() { return 2; }

inline size_t vm_cold_bytes(const bn254_ctx* ctx) {
  int maxb = std::max({ctx->vm_blocks_per_sm[0], ctx->vm_blocks_per_sm[1], ctx->vm_blocks_per_sm[2]});
  return (size_t)kVmColdSlots * 4 * sizeof(uint4) * (size_t)ctx->sms * maxb * kVmNP;
}
template <typename PROG>
void launch_vm(bn254_ctx* ctx, const void* a, const void* b, size_t n, void* out, uint4* cold, cudaStream_t s) {
  size_t want = (n + kVmNP - 1) / kVmNP;
  unsigned grid = (unsigned)std::min<size_t>(want, (size_t)ctx->sms * ctx->vm_blocks_per_sm[vm_prog_index<PROG>()]);
  k_vm<PROG><<<grid, 32 * kVmWarps, vm_smem_bytes<PROG>(), s>>>(a, b, n, out, cold);
}
template <typename PROG>
cudaError_t vm_prepare(bn254_ctx* ctx) {
  cudaError_t e = cudaFuncSetAttribute(k_vm<PROG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)vm_smem_bytes<PROG>());
  if (e != cudaSuccess) return e;
  int nb = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_vm<PROG>, 32 * kVmWarps, vm_smem_bytes<PROG>());
  if (e != cudaSuccess) return e;
  if (nb < 1) return cudaErrorLaunchOutOfResources;
  ctx->vm_blocks_per_sm[vm_prog_index<PROG>()] = nb;
  return cudaSuccess;
}
// device-pointer launches of one context share vm_cold_dev: order them across streams with an event
template <typename PROG>
int run_dev_vm(bn254_ctx* ctx, const void* a, const void* b, size_t n, void* out, cudaStream_t s) {
  if (!ctx) return BN254_ERR_BAD_ARG;
  if (n == 0) return BN254_OK;
  std::lock_guard<std::mutex> lk(ctx->mu);
  CU(cudaSetDevice(ctx->device));
  CU(cudaStreamWaitEvent(s, ctx->vm_dev_done, 0));
  launch_vm<PROG>(ctx, a, b, n, out, ctx->vm_cold_dev, s);
  ctx->launches++;
  CU(cudaGetLastError());
  CU(cudaEventRecord(ctx->vm_dev_done, s));
  return BN254_OK;
}



// fixed-base GT table: entries x^(d << 8w) computed with the generic ladder (valid for any Fp12 base)
int ensure_gt_table(bn254_ctx* ctx, const void* base) {
  std::lock_guard<std::mutex> lk(ctx->mu);
  CU(cudaSetDevice(ctx->device));
  if (ctx->gt_valid && memcmp(ctx->gt_base, base, BN254_GT_BYTES) == 0) return BN254_OK;
  const size_t entries = (size_t)kFixedWindows * kFixedEntries;
  if (!ctx->gt_table) CU(cudaMalloc((void**)&ctx->gt_table, entries * sizeof(Fp12)));
  Slot& s = ctx->slot[0];
  unsigned char* h = reinterpret_cast<unsigned char*>(s.h);
  memcpy(h, base, BN254_GT_BYTES);
  unsigned char* hs = h + 512;
  memset(hs, 0, entries * 32);
  for (int w = 0; w < kFixedWindows; w++)
    for (int d = 1; d <= kFixedEntries; d++) hs[((size_t)w * kFixedEntries + d - 1) * 32 + w] = (unsigned char)d;
  CU(cudaMemcpyAsync(s.d, s.h, 512 + entries * 32, cudaMemcpyHostToDevice, s.stream));
  k_gt_exp<0><<<grid_for(entries), kBlock, kTowerSmem, s.stream>>>(s.d, 0, s.d + 512, entries, ctx->gt_table);
  ctx->launches++;
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(s.stream));
  memcpy(ctx->gt_base, base, BN254_GT_BYTES);
  ctx->gt_valid = true;
  return BN254_OK;
}
int gt_fixed_exp(bn254_ctx* ctx, const void* x1, const void* k, size_t n, void* out) {
  int rc = ensure_gt_table(ctx, x1);
  if (rc) return rc;
  const Fp12* table = ctx->gt_table;
  return run_host(ctx, {k, BN254_SCALAR_BYTES, false}, {nullptr, 0, false}, out, BN254_GT_BYTES, n,
                  [table](const void* a, const void*, size_t c, void* o, cudaStream_t s, uint4*) {
                    k_gt_fixed_exp<<<grid_for(c), kBlock, kTowerSmem, s>>>(table, a, c, o);
                  });
}
// multi-pairing launch: single kernel for small k, split + combine for large k.
// cold != nullptr (host-buffer entry points): products of 2..16 pairs whose total pair count is small are latency-
// bound in the one-thread-per-product kernels (a 1024-message BLS check takes ~15 ms whatever the GPU), so they go
// through the lane-group kernels instead: Miller loop per PAIR (n*k lane groups), product of each k values, final
// exponentiation per product -- ~6 ms for the same batch, same bytes out.
constexpr size_t kVmAutoMax = 16384;
static inline bool use_vm(const bn254_ctx* ctx, size_t n) { return ctx->vm_mode == 1 || (ctx->vm_mode == 0 && n <= kVmAutoMax); }
static cudaError_t ensure_mp_scratch(Slot& sl, size_t need, cudaStream_t s) {
  if (sl.mp_scratch_bytes >= need) return cudaSuccess;
  if (sl.mp_scratch) { cudaStreamSynchronize(s); cudaFree(sl.mp_scratch); sl.mp_scratch = nullptr; sl.mp_scratch_bytes = 0; }
  cudaError_t e = cudaMalloc(&sl.mp_scratch, need);
  if (e == cudaSuccess) sl.mp_scratch_bytes = need;
  return e;
}
template <int MODE>
cudaError_t launch_multi_pair(bn254_ctx* ctx, Slot& sl, uint4* cold, const void* a, const void* b, size_t n, int k, void* o, cudaStream_t s) {
  if (cold && k >= 2 && k <= 2 * kMpChunk && ctx->vm_mode != 2 && use_vm(ctx, n * (size_t)k)) {
    size_t pairs = n * (size_t)k;
    cudaError_t e = ensure_mp_scratch(sl, (pairs + n) * BN254_GT_BYTES, s);
    if (e != cudaSuccess) return e;
    char* ml = static_cast<char*>(sl.mp_scratch);
    void* prod = MODE == 2 ? static_cast<void*>(ml + pairs * BN254_GT_BYTES) : o;
    launch_vm<VmProgMiller>(ctx, a, b, pairs, ml, cold, s);
    k_mp_combine<0><<<grid_for(n), kBlock, kTowerSmem, s>>>(ml, n, k, prod);
    if (MODE >= 1) launch_vm<VmProgFinalExp>(ctx, prod, nullptr, n, prod, cold, s);
    if (MODE == 2) k_gt_is_one<<<grid_for(n), kBlock, 0, s>>>(prod, n, static_cast<uint8_t*>(o));
    return cudaSuccess;
  }
  if (k == 1) { k_multi_pair_c<MODE, 1><<<grid_for(n), kBlock, kTowerSmem, s>>>(a, b, n, o); return cudaSuccess; }
  if (k == 2) { k_multi_pair_c<MODE, 2><<<grid_for(n), kBlock, kTowerSmem, s>>>(a, b, n, o); return cudaSuccess; }
  if (k == 3) { k_multi_pair_c<MODE, 3><<<grid_for(n), kBlock, kTowerSmem, s>>>(a, b, n, o); return cudaSuccess; }
  if (k <= 2 * kMpChunk) { k_multi_pair<MODE><<<grid_for(n), kBlock, kTowerSmem, s>>>(a, b, n, k, o); return cudaSuccess; }
  int nchunks = (k + kMpChunk - 1) / kMpChunk;
  cudaError_t e = ensure_mp_scratch(sl, n * (size_t)nchunks * BN254_GT_BYTES, s);
  if (e != cudaSuccess) return e;
  k_mp_partial<<<dim3(grid_for(n), (unsigned)nchunks), kBlock, kTowerSmem, s>>>(a, b, n, k, nchunks, sl.mp_scratch);
  k_mp_combine<MODE><<<grid_for(n), kBlock, kTowerSmem, s>>>(sl.mp_scratch, n, nchunks, o);
  return cudaSuccess;
}
// One base, n scalars.  Small batches run the GLV kernel on the broadcast base; from kFixedMin scalars on a
// 32 x 255 affine window table of the base is built once (8160 GLV multiplications of d << 8w, cached in the
// context until the base changes) and every scalar costs 32 mixed additions.
constexpr size_t kFixedMin = 4096;
template <typename J, typename A>
int ensure_fixed_table(bn254_ctx* ctx, int g, const void* base) {
  std::lock_guard<std::mutex> lk(ctx->mu);
  CU(cudaSetDevice(ctx->device));
  if (ctx->fixed_valid[g] && memcmp(ctx->fixed_base[g], base, sizeof(A)) == 0) return BN254_OK;
  const size_t entries = (size_t)kFixedWindows * kFixedEntries;
  if (!ctx->fixed_table[g]) CU(cudaMalloc(&ctx->fixed_table[g], entries * sizeof(A)));
  Slot& s = ctx->slot[0];
  unsigned char* h = reinterpret_cast<unsigned char*>(s.h);
  memcpy(h, base, sizeof(A));
  unsigned char* hs = h + 256;
  memset(hs, 0, entries * 32);
  for (int w = 0; w < kFixedWindows; w++)
    for (int d = 1; d <= kFixedEntries; d++) hs[((size_t)w * kFixedEntries + d - 1) * 32 + w] = (unsigned char)d;
  CU(cudaMemcpyAsync(s.d, s.h, 256 + entries * 32, cudaMemcpyHostToDevice, s.stream));
  k_scalar_mul<J, A><<<grid_for(entries), kBlock, 0, s.stream>>>(s.d, 0, s.d + 256, entries, ctx->fixed_table[g]);
  ctx->launches++;
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(s.stream));
  memcpy(ctx->fixed_base[g], base, sizeof(A));
  ctx->fixed_valid[g] = true;
  return BN254_OK;
}

}  // namespace

// n messages (concatenated bytes + n+1 offsets) -> points.  Chunks are sized to one staging slot.
// GT exponentiation in launches of at most one full wave of CTAs, each thread with its own contiguous 4-entry
// (generic) / 16-entry (cyclotomic) table slice in the slot's device scratch (6 KB per thread, <= 350 MB per slot, allocated on first use);
// launches of one call are stream-ordered, so consecutive waves reuse the same slices.
template <int CYCLO>
static void launch_gt_exp(bn254_ctx* ctx, uint4* cold, const void* x, size_t stride, const void* k, size_t n, void* o, cudaStream_t s) {
  Slot& sl = ctx->slot[cold == ctx->slot[1].vm_cold ? 1 : 0];
  const size_t wave = (size_t)ctx->sms * BN254_MIN_BLOCKS * kBlock;
  if (ensure_mp_scratch(sl, std::min(n, wave) * (CYCLO ? kGtCycloTable : 4) * sizeof(Fp12), s) != cudaSuccess) return;  // run_host reports cudaGetLastError()
  for (size_t off = 0; off < n; off += wave) {
    size_t c = std::min(wave, n - off);
    k_gt_exp<CYCLO><<<grid_for(c), kBlock, kTowerSmem, s>>>(static_cast<const char*>(x) + off * stride * BN254_GT_BYTES, stride,
                                                        static_cast<const char*>(k) + off * BN254_SCALAR_BYTES, c,
                                                        static_cast<char*>(o) + off * BN254_GT_BYTES, static_cast<Fp12*>(sl.mp_scratch));
  }
}

template <int G>
int hash_to_curve_host(bn254_ctx* ctx, const uint8_t* msgs, const uint64_t* offsets, size_t n, const uint8_t* dst, size_t dst_len, void* out) {
  if (!ctx) return BN254_ERR_BAD_ARG;
  if (dst_len > 255) return fail(ctx, BN254_ERR_BAD_ARG, "hash-to-curve: domain separation tag longer than 255 bytes");
  if (n == 0) return BN254_OK;
  if (!msgs && offsets && offsets[n] != offsets[0]) return fail(ctx, BN254_ERR_BAD_ARG, "null pointer");
  if (!offsets || !out || (dst_len && !dst)) return fail(ctx, BN254_ERR_BAD_ARG, "null pointer");
  const size_t out_item = G == 1 ? BN254_G1_BYTES : BN254_G2_BYTES;
  std::lock_guard<std::mutex> lk(ctx->mu);
  CU(cudaSetDevice(ctx->device));
  // Two slots in flight: chunk i's kernel runs while chunk i-1's points are copied back and handed to the caller.
  // Chunks are capped so that a large batch splits into >= 4 of them (a whole wave of CTAs each at the least).
  const size_t cap = std::min<size_t>(kMaxChunkItems, std::max<size_t>((n + 3) / 4, 148 * 3 * kBlock));
  struct Pending { size_t done = 0, c = 0, o_out = 0; bool live = false; } pend[2];
  auto drain = [&](int i) -> int {
    if (!pend[i].live) return BN254_OK;
    CU(cudaStreamSynchronize(ctx->slot[i].stream));
    memcpy(static_cast<char*>(out) + pend[i].done * out_item, ctx->slot[i].h + pend[i].o_out, out_item * pend[i].c);
    pend[i].live = false;
    return BN254_OK;
  };
  size_t done = 0;
  for (int it = 0; done < n; it ^= 1) {
    if (int rc = drain(it)) return rc;
    Slot& s = ctx->slot[it];
    // largest c with  256 (dst) + 8 (c + 1) + bytes + out_item * c  <=  slot
    size_t c = 0, bytes = 0;
    while (done + c < n && c < cap) {
      size_t len = (size_t)(offsets[done + c + 1] - offsets[done + c]);
      if (512 + 8 * (c + 2) + bytes + len + out_item * (c + 1) + 512 > ctx->slot_bytes) break;
      bytes += len; c++;
    }
    if (c == 0) { drain(it ^ 1); return fail(ctx, BN254_ERR_BAD_ARG, "message too large for staging"); }
    unsigned char* h = reinterpret_cast<unsigned char*>(s.h);
    memset(h, 0, 256);
    if (dst_len) memcpy(h, dst, dst_len);
    uint64_t* ho = reinterpret_cast<uint64_t*>(h + 256);
    for (size_t j = 0; j <= c; j++) ho[j] = offsets[done + j] - offsets[done];
    size_t o_msg = (256 + 8 * (c + 1) + 255) & ~size_t(255);
    if (bytes) memcpy(h + o_msg, msgs + offsets[done], bytes);
    size_t o_out = (o_msg + bytes + 255) & ~size_t(255);
    CU(cudaMemcpyAsync(s.d, s.h, o_out, cudaMemcpyHostToDevice, s.stream));
    const uint8_t* d = reinterpret_cast<const uint8_t*>(s.d);
    // no tower scratch in this kernel: launched without dynamic shared memory, G1 (90 registers) runs 5 CTAs per SM
    k_hash_to_curve<G><<<grid_for(c), kBlock, 0, s.stream>>>(d + o_msg, reinterpret_cast<const uint64_t*>(d + 256), c, d, (uint32_t)dst_len,
                                                                     s.d + o_out);
    ctx->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(s.h + o_out, s.d + o_out, out_item * c, cudaMemcpyDeviceToHost, s.stream));
    pend[it].done = done; pend[it].c = c; pend[it].o_out = o_out; pend[it].live = true;
    done += c;
  }
  if (int rc = drain(0)) return rc;
  if (int rc = drain(1)) return rc;
  return BN254_OK;
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
#ifdef BN254_SMEM_SCRATCH
  {
    const void* tower_kernels[] = {(const void*)k_pair, (const void*)k_gt_fixed_exp, (const void*)k_g2_lines, (const void*)k_miller_lines, (const void*)k_multi_pair_c<0, 1>, (const void*)k_multi_pair_c<1, 1>, (const void*)k_multi_pair_c<2, 1>,
                                   (const void*)k_multi_pair_c<0, 2>, (const void*)k_multi_pair_c<1, 2>, (const void*)k_multi_pair_c<2, 2>,
                                   (const void*)k_multi_pair_c<0, 3>, (const void*)k_multi_pair_c<1, 3>, (const void*)k_multi_pair_c<2, 3>, (const void*)k_multi_pair<0>, (const void*)k_multi_pair<1>, (const void*)k_multi_pair<2>,
                                   (const void*)k_mp_partial, (const void*)k_mp_combine<0>, (const void*)k_mp_combine<1>, (const void*)k_mp_combine<2>,
                                   (const void*)k_final_exp, (const void*)k_check2_fixed_g1, (const void*)k_hash_to_curve<1>, (const void*)k_hash_to_curve<2>, (const void*)k_gt_exp<0>, (const void*)k_gt_exp<1>, (const void*)k_gt_mul<0>, (const void*)k_gt_mul<1>};
    for (const void* k : tower_kernels)
      if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTowerSmem) != cudaSuccess) { bn254_ctx_destroy(ctx); return BN254_ERR_CUDA; }
  }
#endif
  const char* impl = getenv("BN254_IMPL");
  ctx->vm_mode = !impl ? 0 : (std::string(impl) == "vm" ? 1 : (std::string(impl) == "thread" ? 2 : 0));
  ctx->sms = prop.multiProcessorCount;
  if (vm_prepare<VmProgPair>(ctx) != cudaSuccess || vm_prepare<VmProgMiller>(ctx) != cudaSuccess ||
      vm_prepare<VmProgFinalExp>(ctx) != cudaSuccess) { bn254_ctx_destroy(ctx); return BN254_ERR_CUDA; }
  size_t cb = vm_cold_bytes(ctx);
  if (cudaMalloc((void**)&ctx->vm_cold_dev, cb) != cudaSuccess || cudaMalloc((void**)&ctx->slot[0].vm_cold, cb) != cudaSuccess ||
      cudaMalloc((void**)&ctx->slot[1].vm_cold, cb) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->vm_dev_done, cudaEventDisableTiming) != cudaSuccess) { bn254_ctx_destroy(ctx); return BN254_ERR_OOM; }
  *out = ctx;
  return BN254_OK;
}

void bn254_ctx_destroy(bn254_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  for (int i = 0; i < 2; i++) {
    Slot& s = ctx->slot[i];
    if (s.stream) { cudaStreamSynchronize(s.stream); cudaStreamDestroy(s.stream); }
    if (s.h) cudaFreeHost(s.h);
    if (s.d) cudaFree(s.d);
    if (s.vm_cold) cudaFree(s.vm_cold);
    if (s.mp_scratch) cudaFree(s.mp_scratch);
  }
  if (ctx->dev_slot.mp_scratch) { cudaFree(ctx->dev_slot.mp_scratch); }
  if (ctx->vm_cold_dev) cudaFree(ctx->vm_cold_dev);
  for (int g = 0; g < 2; g++) if (ctx->fixed_table[g]) cudaFree(ctx->fixed_table[g]);
  if (ctx->gt_table) cudaFree(ctx->gt_table);
  if (ctx->vm_dev_done) cudaEventDestroy(ctx->vm_dev_done);
  delete ctx;
}

const char* bn254_last_error(bn254_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
uint64_t bn254_launch_count(bn254_ctx* ctx) { return ctx ? ctx->launches : 0; }

void* bn254_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
  return p;
}
void bn254_host_free(void* p) { if (p) cudaFreeHost(p); }

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
  if (ctx && use_vm(ctx, n)) return run_dev_vm<VmProgPair>(ctx, dP, dQ, n, d_out, (cudaStream_t)stream);
  return run_dev(ctx, n, [&] { k_pair<<<grid_for(n), kBlock, kTowerSmem, (cudaStream_t)stream>>>(dP, dQ, n, d_out); });
}
int bn254_pair_batch(bn254_ctx* ctx, const void* P, const void* Q, size_t n, void* out) {
  return run_host(ctx, {P, BN254_G1_BYTES, false}, {Q, BN254_G2_BYTES, false}, out, BN254_GT_BYTES, n,
                  [ctx](const void* a, const void* b, size_t c, void* o, cudaStream_t s, uint4* cold) {
                    if (use_vm(ctx, c)) launch_vm<VmProgPair>(ctx, a, b, c, o, cold, s);
                    else k_pair<<<grid_for(c), kBlock, kTowerSmem, s>>>(a, b, c, o);
                  });
}
#define MULTI_PAIR_ENTRY(name, MODE, OUT_BYTES, OUT_T)                                                                     \
  int name##_dev(bn254_ctx* ctx, const void* dP, const void* dQ, size_t n, size_t k, OUT_T* d_out, void* stream) {         \
    if (k == 0 || k > (1u << 20)) return fail(ctx, BN254_ERR_INVALID_SIZES, "invalid inputs sizes");                       \
    if (ctx && use_vm(ctx, n) && k == 1 && MODE == 0) return run_dev_vm<VmProgMiller>(ctx, dP, dQ, n, d_out, (cudaStream_t)stream); \
    if (ctx && use_vm(ctx, n) && k == 1 && MODE == 1) return run_dev_vm<VmProgPair>(ctx, dP, dQ, n, d_out, (cudaStream_t)stream);   \
    return run_dev(ctx, n, [&] { cudaError_t e_ = launch_multi_pair<MODE>(ctx, ctx->dev_slot, nullptr, dP, dQ, n, (int)k, d_out, (cudaStream_t)stream); (void)e_; }); \
  }                                                                                                                        \
  int name(bn254_ctx* ctx, const void* P, const void* Q, size_t n, size_t k, OUT_T* out) {                                 \
    if (k == 0 || k > (1u << 20)) return fail(ctx, BN254_ERR_INVALID_SIZES, "invalid inputs sizes");                       \
    int kk = (int)k;                                                                                                       \
    return run_host(ctx, {P, BN254_G1_BYTES * k, false}, {Q, BN254_G2_BYTES * k, false}, out, OUT_BYTES, n,                \
                    [kk, ctx](const void* a, const void* b, size_t c, void* o, cudaStream_t s, uint4* cold) {                  \
                      if (use_vm(ctx, c) && kk == 1 && MODE == 0) launch_vm<VmProgMiller>(ctx, a, b, c, o, cold, s);          \
                      else if (use_vm(ctx, c) && kk == 1 && MODE == 1) launch_vm<VmProgPair>(ctx, a, b, c, o, cold, s);       \
                      else launch_multi_pair<MODE>(ctx, ctx->slot[cold == ctx->slot[1].vm_cold ? 1 : 0], cold, a, b, c, kk, o, s);   \
                    });                                                                                                    \
  }
MULTI_PAIR_ENTRY(bn254_miller_loop_batch, 0, BN254_GT_BYTES, void)
MULTI_PAIR_ENTRY(bn254_multi_pair_batch, 1, BN254_GT_BYTES, void)
MULTI_PAIR_ENTRY(bn254_pairing_check_batch, 2, 1, uint8_t)

// ---- line tables --------------------------------------------------------------------------------------
struct bn254_lines {
  bn254_ctx* ctx;
  size_t m;
  Fp2* table;      // m x kLinesPerPoint x 3 Fp2 on the device
  uint8_t* qskip;  // m flags: point at infinity
};
int bn254_g2_lines_create(bn254_ctx* ctx, const void* Q, size_t m, bn254_lines** out) {
  if (!ctx || !Q || !out || m == 0) return fail(ctx, BN254_ERR_BAD_ARG, "bad lines arguments");
  std::lock_guard<std::mutex> lk(ctx->mu);
  CU(cudaSetDevice(ctx->device));
  if (m * BN254_G2_BYTES > ctx->slot_bytes) return fail(ctx, BN254_ERR_BAD_ARG, "too many points for one table");
  bn254_lines* L = new bn254_lines{ctx, m, nullptr, nullptr};
  size_t bytes = m * (size_t)kLinesPerPoint * 3 * sizeof(Fp2);
  if (cudaMalloc((void**)&L->table, bytes) != cudaSuccess || cudaMalloc((void**)&L->qskip, m) != cudaSuccess) {
    if (L->table) cudaFree(L->table);
    delete L;
    return fail(ctx, BN254_ERR_OOM, "line table allocation");
  }
  Slot& s = ctx->slot[0];
  memcpy(s.h, Q, m * BN254_G2_BYTES);
  CU(cudaMemcpyAsync(s.d, s.h, m * BN254_G2_BYTES, cudaMemcpyHostToDevice, s.stream));
  k_g2_lines<<<grid_for(m), kBlock, kTowerSmem, s.stream>>>(s.d, m, L->table, L->qskip);
  ctx->launches++;
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(s.stream));
  *out = L;
  return BN254_OK;
}
void bn254_g2_lines_destroy(bn254_lines* L) {
  if (!L) return;
  cudaSetDevice(L->ctx->device);
  cudaFree(L->table);
  cudaFree(L->qskip);
  delete L;
}
size_t bn254_g2_lines_count(const bn254_lines* L) { return L ? L->m : 0; }
// out[i] = FinalExponentiation(prod_j Miller(P[i*m + j], Q_j)) for the m table points: n products of m pairs
int bn254_multi_pair_lines_batch(bn254_ctx* ctx, const void* P, const bn254_lines* L, size_t n, void* out) {
  if (!ctx || !L || L->ctx != ctx) return fail(ctx, BN254_ERR_BAD_ARG, "line table belongs to another context");
  const int m = (int)L->m;
  const int nchunks = (m + kMpChunk - 1) / kMpChunk;
  const Fp2* table = L->table;
  const uint8_t* qskip = L->qskip;
  return run_host(ctx, {P, BN254_G1_BYTES * (size_t)m, false}, {nullptr, 0, false}, out, BN254_GT_BYTES, n,
                  [=](const void* a, const void*, size_t c, void* o, cudaStream_t s, uint4* cold) {
                    Slot& sl = ctx->slot[cold == ctx->slot[1].vm_cold ? 1 : 0];
                    size_t need = c * (size_t)nchunks * BN254_GT_BYTES;
                    if (sl.mp_scratch_bytes < need) {
                      if (sl.mp_scratch) { cudaStreamSynchronize(s); cudaFree(sl.mp_scratch); sl.mp_scratch = nullptr; sl.mp_scratch_bytes = 0; }
                      if (cudaMalloc(&sl.mp_scratch, need) != cudaSuccess) return;
                      sl.mp_scratch_bytes = need;
                    }
                    dim3 grid(grid_for(c), (unsigned)nchunks);
                    k_miller_lines<<<grid, kBlock, kTowerSmem, s>>>(a, table, qskip, c, m, nchunks, sl.mp_scratch);
                    k_mp_combine<1><<<grid_for(c), kBlock, kTowerSmem, s>>>(sl.mp_scratch, c, nchunks, o);
                  });
}

int bn254_final_exp_batch_dev(bn254_ctx* ctx, const void* d_in, size_t n, void* d_out, void* stream) {
  if (ctx && use_vm(ctx, n)) return run_dev_vm<VmProgFinalExp>(ctx, d_in, nullptr, n, d_out, (cudaStream_t)stream);
  return run_dev(ctx, n, [&] { k_final_exp<<<grid_for(n), kBlock, kTowerSmem, (cudaStream_t)stream>>>(d_in, n, d_out); });
}
int bn254_final_exp_batch(bn254_ctx* ctx, const void* in, size_t n, void* out) {
  return run_host(ctx, {in, BN254_GT_BYTES, false}, {nullptr, 0, false}, out, BN254_GT_BYTES, n,
                  [ctx](const void* a, const void*, size_t c, void* o, cudaStream_t s, uint4* cold) {
                    if (use_vm(ctx, c)) launch_vm<VmProgFinalExp>(ctx, a, nullptr, c, o, cold, s);
                    else k_final_exp<<<grid_for(c), kBlock, kTowerSmem, s>>>(a, c, o);
                  });
}

// ---- scalar multiplication --------------------------------------------------------------------
int bn254_g1_mul_batch_dev(bn254_ctx* ctx, const void* d_base, size_t stride, const void* d_s, size_t n, void* d_out, void* stream) {
  return run_dev(ctx, n, [&] { k_scalar_mul<G1Jac, G1Aff><<<grid_for(n), kBlock, 0, (cudaStream_t)stream>>>(d_base, stride, d_s, n, d_out); });
}
int bn254_g2_mul_batch_dev(bn254_ctx* ctx, const void* d_base, size_t stride, const void* d_s, size_t n, void* d_out, void* stream) {
  return run_dev(ctx, n, [&] { k_scalar_mul<G2Jac, G2Aff><<<grid_for(n), kBlock, 0, (cudaStream_t)stream>>>(d_base, stride, d_s, n, d_out); });
}
#define MUL_ENTRY(name, J, A, BYTES, BCAST)                                                                           \
  int name(bn254_ctx* ctx, const void* base, const void* scalars, size_t n, void* out) {                              \
    return run_host(ctx, {base, BYTES, BCAST}, {scalars, BN254_SCALAR_BYTES, false}, out, BYTES, n,                   \
                    [](const void* a, const void* b, size_t c, void* o, cudaStream_t s, uint4*) {                             \
                      k_scalar_mul<J, A><<<grid_for(c), kBlock, 0, s>>>(a, BCAST ? 0 : 1, b, c, o);                   \
                    });                                                                                               \
  }
MUL_ENTRY(bn254_g1_mul_batch, G1Jac, G1Aff, BN254_G1_BYTES, false)
MUL_ENTRY(bn254_g2_mul_batch, G2Jac, G2Aff, BN254_G2_BYTES, false)
#define MUL_BASE_ENTRY(name, J, A, BYTES, G)                                                                          \
  int name(bn254_ctx* ctx, const void* base, const void* scalars, size_t n, void* out) {                              \
    if (!ctx || !base) return BN254_ERR_BAD_ARG;                                                                      \
    if (n < kFixedMin)                                                                                                \
      return run_host(ctx, {base, BYTES, true}, {scalars, BN254_SCALAR_BYTES, false}, out, BYTES, n,                  \
                      [](const void* a, const void* b, size_t c, void* o, cudaStream_t s, uint4*) {                   \
                        k_scalar_mul<J, A><<<grid_for(c), kBlock, 0, s>>>(a, 0, b, c, o);                             \
                      });                                                                                             \
    int rc = ensure_fixed_table<J, A>(ctx, G, base);                                                                  \
    if (rc) return rc;                                                                                                \
    const A* table = static_cast<const A*>(ctx->fixed_table[G]);                                                      \
    return run_host(ctx, {base, BYTES, true}, {scalars, BN254_SCALAR_BYTES, false}, out, BYTES, n,                    \
                    [table](const void*, const void* b, size_t c, void* o, cudaStream_t s, uint4*) {                  \
                      k_fixed_mul<J, A><<<grid_for(c), kBlock, 0, s>>>(table, b, c, o);                               \
                    });                                                                                               \
  }
MUL_BASE_ENTRY(bn254_g1_mul_base_batch, G1Jac, G1Aff, BN254_G1_BYTES, 0)
MUL_BASE_ENTRY(bn254_g2_mul_base_batch, G2Jac, G2Aff, BN254_G2_BYTES, 1)

// ---- subset sums and segment sums -----------------------------------------------------------------
#define SUBSET_SUM_ENTRY(name, J, A, BYTES)                                                                              \
  int name(bn254_ctx* ctx, const void* U, size_t m, const void* sel, size_t n, void* out) {                              \
    if (!ctx || !U || m == 0 || (m + 1) * BYTES > 200 * 1024) return fail(ctx, BN254_ERR_BAD_ARG, "bad subset-sum arguments"); \
    size_t smem = (m + 1) * BYTES;                                                                                       \
    { std::lock_guard<std::mutex> lk(ctx->mu); CU(cudaSetDevice(ctx->device));                                          \
      CU(cudaFuncSetAttribute(k_subset_sum<J, A>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); }           \
    int mm = (int)m;                                                                                                     \
    return run_host(ctx, {U, (m + 1) * BYTES, true}, {sel, (m + 7) / 8, false}, out, BYTES, n,                           \
                    [mm, smem](const void* a, const void* b, size_t c, void* o, cudaStream_t s, uint4*) {                \
                      k_subset_sum<J, A><<<grid_for(c), kBlock, smem, s>>>(static_cast<const A*>(a), mm,                 \
                                                                            static_cast<const uint8_t*>(b), c, o);        \
                    });                                                                                                  \
  }
SUBSET_SUM_ENTRY(bn254_g1_subset_sum_batch, G1Jac, G1Aff, BN254_G1_BYTES)
SUBSET_SUM_ENTRY(bn254_g2_subset_sum_batch, G2Jac, G2Aff, BN254_G2_BYTES)

// out[g] = sum of points[g*len .. g*len+len): passes of 32-way partial sums, all on the device
#define SEGMENT_SUM_ENTRY(name, J, A, BYTES)                                                                             \
  int name(bn254_ctx* ctx, const void* pts, size_t groups, size_t len, void* out) {                                      \
    if (!ctx || !pts || !out || len == 0 || len > (1u << 24)) return fail(ctx, BN254_ERR_BAD_ARG, "bad segment-sum arguments"); \
    if (groups == 0) return BN254_OK;                                                                                    \
    std::lock_guard<std::mutex> lk(ctx->mu);                                                                             \
    CU(cudaSetDevice(ctx->device));                                                                                      \
    Slot& s = ctx->slot[0];                                                                                              \
    const int chunk = 32;                                                                                                \
    size_t per_group = len * BYTES + ((len + chunk - 1) / chunk) * BYTES + 512;                                          \
    size_t gmax = (ctx->slot_bytes - 4096) / per_group;                                                                  \
    if (gmax == 0) return fail(ctx, BN254_ERR_BAD_ARG, "segment too long for staging");                                  \
    for (size_t g0 = 0; g0 < groups; g0 += gmax) {                                                                       \
      size_t gc = std::min(gmax, groups - g0);                                                                           \
      size_t in_bytes = gc * len * BYTES;                                                                                \
      memcpy(s.h, static_cast<const char*>(pts) + g0 * len * BYTES, in_bytes);                                           \
      CU(cudaMemcpyAsync(s.d, s.h, in_bytes, cudaMemcpyHostToDevice, s.stream));                                         \
      char* cur = s.d;                                                                                                   \
      char* nxt = s.d + ((in_bytes + 255) & ~size_t(255));                                                               \
      size_t cur_len = len;                                                                                              \
      while (true) {                                                                                                     \
        int nch = (int)((cur_len + chunk - 1) / chunk);                                                                  \
        k_segment_sum<J, A><<<grid_for(gc * (size_t)nch), kBlock, 0, s.stream>>>(cur, gc, (int)cur_len, chunk, nxt);     \
        ctx->launches++;                                                                                                 \
        CU(cudaGetLastError());                                                                                          \
        std::swap(cur, nxt);                                                                                             \
        cur_len = (size_t)nch;                                                                                           \
        if (nch == 1) break;                                                                                             \
      }                                                                                                                  \
      CU(cudaMemcpyAsync(s.h, cur, gc * BYTES, cudaMemcpyDeviceToHost, s.stream));                                       \
      CU(cudaStreamSynchronize(s.stream));                                                                               \
      memcpy(static_cast<char*>(out) + g0 * BYTES, s.h, gc * BYTES);                                                     \
    }                                                                                                                    \
    return BN254_OK;                                                                                                     \
  }
SEGMENT_SUM_ENTRY(bn254_g1_sum_batch, G1Jac, G1Aff, BN254_G1_BYTES)
SEGMENT_SUM_ENTRY(bn254_g2_sum_batch, G2Jac, G2Aff, BN254_G2_BYTES)

int bn254_g1_add_batch(bn254_ctx* ctx, const void* a, const void* b, size_t n, void* out) {
  return run_host(ctx, {a, BN254_G1_BYTES, false}, {b, BN254_G1_BYTES, false}, out, BN254_G1_BYTES, n,
                  [](const void* x, const void* y, size_t c, void* o, cudaStream_t s, uint4*) { k_aff_add<G1Jac, G1Aff><<<grid_for(c), kBlock, 0, s>>>(x, y, c, o); });
}
int bn254_g2_add_batch(bn254_ctx* ctx, const void* a, const void* b, size_t n, void* out) {
  return run_host(ctx, {a, BN254_G2_BYTES, false}, {b, BN254_G2_BYTES, false}, out, BN254_G2_BYTES, n,
                  [](const void* x, const void* y, size_t c, void* o, cudaStream_t s, uint4*) { k_aff_add<G2Jac, G2Aff><<<grid_for(c), kBlock, 0, s>>>(x, y, c, o); });
}

// ---- GT ---------------------------------------------------------------------------------------
int bn254_gt_exp_batch_dev(bn254_ctx* ctx, const void* d_x, size_t stride, const void* d_k, size_t n, void* d_out, void* stream) {
  return run_dev(ctx, n, [&] { k_gt_exp<0><<<grid_for(n), kBlock, kTowerSmem, (cudaStream_t)stream>>>(d_x, stride, d_k, n, d_out); });
}
int bn254_gt_exp_batch(bn254_ctx* ctx, const void* x, const void* k, size_t n, void* out) {
  return run_host(ctx, {x, BN254_GT_BYTES, false}, {k, BN254_SCALAR_BYTES, false}, out, BN254_GT_BYTES, n,
                  [ctx](const void* a, const void* b, size_t c, void* o, cudaStream_t s, uint4* cold) {
                    // one wave or less: table in the slot scratch; larger chunks keep ONE launch with the table on the stack
                    // (wave-sized launches in series measured 14 % slower here than the single launch)
                    if (c <= (size_t)ctx->sms * BN254_MIN_BLOCKS * kBlock) launch_gt_exp<0>(ctx, cold, a, 1, b, c, o, s);
                    else k_gt_exp<0><<<grid_for(c), kBlock, kTowerSmem, s>>>(a, 1, b, c, o);
                  });
}
int bn254_gt_exp_base_batch(bn254_ctx* ctx, const void* x1, const void* k, size_t n, void* out) {
  if (ctx && x1 && n >= kFixedMin) return gt_fixed_exp(ctx, x1, k, n, out);
  return run_host(ctx, {x1, BN254_GT_BYTES, true}, {k, BN254_SCALAR_BYTES, false}, out, BN254_GT_BYTES, n,
                  [ctx](const void* a, const void* b, size_t c, void* o, cudaStream_t s, uint4* cold) {
                    if (c <= (size_t)ctx->sms * BN254_MIN_BLOCKS * kBlock) launch_gt_exp<0>(ctx, cold, a, 0, b, c, o, s);
                    else k_gt_exp<0><<<grid_for(c), kBlock, kTowerSmem, s>>>(a, 0, b, c, o);
                  });
}
int bn254_gt_cyclo_exp_batch(bn254_ctx* ctx, const void* x, const void* k, size_t n, void* out) {
  return run_host(ctx, {x, BN254_GT_BYTES, false}, {k, BN254_SCALAR_BYTES, false}, out, BN254_GT_BYTES, n,
                  [ctx](const void* a, const void* b, size_t c, void* o, cudaStream_t s, uint4* cold) { launch_gt_exp<1>(ctx, cold, a, 1, b, c, o, s); });
}
int bn254_gt_cyclo_exp_base_batch(bn254_ctx* ctx, const void* x1, const void* k, size_t n, void* out) {
  if (ctx && x1 && n >= kFixedMin) return gt_fixed_exp(ctx, x1, k, n, out);
  return run_host(ctx, {x1, BN254_GT_BYTES, true}, {k, BN254_SCALAR_BYTES, false}, out, BN254_GT_BYTES, n,
                  [ctx](const void* a, const void* b, size_t c, void* o, cudaStream_t s, uint4* cold) { launch_gt_exp<1>(ctx, cold, a, 0, b, c, o, s); });
}
int bn254_gt_mul_batch(bn254_ctx* ctx, const void* a, const void* b, size_t n, void* out) {
  return run_host(ctx, {a, BN254_GT_BYTES, false}, {b, BN254_GT_BYTES, false}, out, BN254_GT_BYTES, n,
                  [](const void* x, const void* y, size_t c, void* o, cudaStream_t s, uint4*) { k_gt_mul<0><<<grid_for(c), kBlock, kTowerSmem, s>>>(x, y, c, o); });
}
int bn254_gt_div_batch(bn254_ctx* ctx, const void* a, const void* b, size_t n, void* out) {
  return run_host(ctx, {a, BN254_GT_BYTES, false}, {b, BN254_GT_BYTES, false}, out, BN254_GT_BYTES, n,
                  [](const void* x, const void* y, size_t c, void* o, cudaStream_t s, uint4*) { k_gt_mul<1><<<grid_for(c), kBlock, kTowerSmem, s>>>(x, y, c, o); });
}
int bn254_fp_mul_batch(bn254_ctx* ctx, const void* a, const void* b, size_t n, void* out) {
  return run_host(ctx, {a, 32, false}, {b, 32, false}, out, 32, n,
                  [](const void* x, const void* y, size_t c, void* o, cudaStream_t s, uint4*) { k_fp_mul<<<grid_for(c), kBlock, 0, s>>>(x, y, c, o); });
}

// ---- BLS-shaped check: two G1 points fixed for the batch ---------------------------------------------------------
int bn254_pairing_check2_fixed_g1_batch(bn254_ctx* ctx, const void* P01, const void* Q0, const void* Q1, size_t n, uint8_t* ok) {
  if (n && (!P01 || !Q0 || !Q1)) return fail(ctx, BN254_ERR_BAD_ARG, "null pointer");
  Operand in[3] = {{P01, 2 * BN254_G1_BYTES, true}, {Q0, BN254_G2_BYTES, false}, {Q1, BN254_G2_BYTES, false}};
  return run_host_n(ctx, in, 3, ok, 1, n, [ctx](const void* const* d, size_t c, void* o, cudaStream_t s, uint4* cold) {
    Slot& sl = ctx->slot[cold == ctx->slot[1].vm_cold ? 1 : 0];
    // small batch: lane-group kernels (see launch_multi_pair); falls through to the thread kernel if the scratch cannot grow
    if (ctx->vm_mode != 2 && use_vm(ctx, 2 * c) &&
        ensure_mp_scratch(sl, (size_t)(2 * c) * (BN254_G1_BYTES + BN254_G2_BYTES + BN254_GT_BYTES) + c * BN254_GT_BYTES, s) == cudaSuccess) {
      char* base = static_cast<char*>(sl.mp_scratch);
      char* Pp = base; char* Qp = Pp + 2 * c * BN254_G1_BYTES; char* ml = Qp + 2 * c * BN254_G2_BYTES; char* prod = ml + 2 * c * BN254_GT_BYTES;
      k_pack_check2<<<grid_for(c), kBlock, 0, s>>>(d[0], d[1], d[2], c, Pp, Qp);
      launch_vm<VmProgMiller>(ctx, Pp, Qp, 2 * c, ml, cold, s);
      k_mp_combine<0><<<grid_for(c), kBlock, kTowerSmem, s>>>(ml, c, 2, prod);
      launch_vm<VmProgFinalExp>(ctx, prod, nullptr, c, prod, cold, s);
      k_gt_is_one<<<grid_for(c), kBlock, 0, s>>>(prod, c, static_cast<uint8_t*>(o));
      return;
    }
    k_check2_fixed_g1<<<grid_for(c), kBlock, kTowerSmem, s>>>(d[0], d[1], d[2], c, static_cast<uint8_t*>(o));
  });
}

// ---- hash-to-curve (gnark bn254.HashToG1 / HashToG2) -----------------------------------------------------------
int bn254_hash_to_g1_batch(bn254_ctx* ctx, const uint8_t* msgs, const uint64_t* offsets, size_t n, const uint8_t* dst, size_t dst_len, void* out) {
  return hash_to_curve_host<1>(ctx, msgs, offsets, n, dst, dst_len, out);
}
int bn254_hash_to_g2_batch(bn254_ctx* ctx, const uint8_t* msgs, const uint64_t* offsets, size_t n, const uint8_t* dst, size_t dst_len, void* out) {
  return hash_to_curve_host<2>(ctx, msgs, offsets, n, dst, dst_len, out);
}
}  // extern "C"
