// Group family, additions: gnark Add semantics, Waters-hash subset sums (per bit and byte-window tables), segment sums, negation.
#include "kcommon.cuh"
#include "curve.cuh"

namespace bn254 {
namespace {
template <typename J, typename A>
__global__ void __launch_bounds__(kBlock, BN254_MIN_BLOCKS) k_aff_add(const void* a, const void* b, size_t n, void* out) {
  cta_lockstep_set(false);  // no barriers in this kernel; the flag is read by the shared field routines
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < n;
  A x, y, r;
  if (live) { load_struct(x, a, i); load_struct(y, b, i); }
  else { f_set_zero(x.x); f_set_zero(x.y); y = x; }
  // gnark Add semantics with the CTA's shared inversion: infinity operands pass the other one through (z = 1 is
  // inverted, harmlessly), everything else is one mixed addition and a shared normalisation
  J t;
  if (aff_is_inf(x)) { t.x = y.x; t.y = y.y; f_set_one(t.z); if (aff_is_inf(y)) f_set_zero(t.z); }
  else { t.x = x.x; t.y = x.y; f_set_one(t.z); if (!aff_is_inf(y)) jac_add_aff(t, t, y); }
  jac_to_aff_inv<InvCta>(r, t);
  if (live) store_struct(out, i, r);
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
  const bool live = i < n;
  const uint8_t* bits = sel + (live ? i : 0) * (size_t)((m + 7) / 8);
  J acc;
  if (aff_is_inf(su[0])) { f_set_zero(acc.x); f_set_zero(acc.y); f_set_zero(acc.z); }
  else { acc.x = su[0].x; acc.y = su[0].y; f_set_one(acc.z); }
  for (int j = 0; j < (live ? m : 0); j++) {
    if ((bits[j >> 3] >> (7 - (j & 7))) & 1) {
      A e = su[j + 1];
      if (!aff_is_inf(e)) jac_add_aff(acc, acc, e);
    }
  }
  A r;
  jac_to_aff_inv<InvCta>(r, acc);  // one inversion per CTA (threads past the end take part with the base point)
  if (live) store_struct(out, i, r);
}
// Byte-window form of the same subset sum for large batches: table[b * 256 + v] = sum of the points U[1 + 8b + i] whose
// bit (7 - i) is set in v (affine; v = 0 -> infinity), built once per call by 256 threads per selector byte; an
// identity then costs ceil(m/8) mixed additions instead of ~m/2 (Waters05, m = 256: 32 instead of ~128).
template <typename J, typename A>
__global__ void __launch_bounds__(kBlock, BN254_MIN_BLOCKS) k_subset_table(const A* U, int m, A* table) {
  cta_lockstep_set(false);
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  int nbytes = (m + 7) / 8;
  if (t >= nbytes * 256) return;
  int b = t >> 8, v = t & 255;
  J acc; f_set_zero(acc.x); f_set_zero(acc.y); f_set_zero(acc.z);
  for (int i = 0; i < 8; i++) {
    int j = 8 * b + i;
    if (j < m && ((v >> (7 - i)) & 1)) {
      A e = U[j + 1];
      if (!aff_is_inf(e)) jac_add_aff(acc, acc, e);
    }
  }
  A r;
  jac_to_aff(r, acc);
  table[t] = r;
}
template <typename J, typename A>
__global__ void __launch_bounds__(kBlock, BN254_MIN_BLOCKS) k_subset_sum_tab(const A* U, const A* __restrict__ table, int m, const uint8_t* sel, size_t n, void* out) {
  cta_lockstep_set(false);
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < n;
  int nbytes = (m + 7) / 8;
  const uint8_t* bits = sel + (live ? i : 0) * (size_t)nbytes;
  J acc;
  A u0 = U[0];
  if (aff_is_inf(u0)) { f_set_zero(acc.x); f_set_zero(acc.y); f_set_zero(acc.z); }
  else { acc.x = u0.x; acc.y = u0.y; f_set_one(acc.z); }
  for (int b = 0; b < (live ? nbytes : 0); b++) {
    int v = bits[b];
    if (v) {
      A e; load_struct(e, table, (size_t)b * 256 + v);
      if (!aff_is_inf(e)) jac_add_aff(acc, acc, e);
    }
  }
  A r;
  jac_to_aff_inv<InvCta>(r, acc);  // one inversion per CTA
  if (live) store_struct(out, i, r);
}
// out[g] = sum of the `len` consecutive points of group g, processed as ceil(len/32)-way partial sums per pass
template <typename J, typename A>
__global__ void __launch_bounds__(kBlock, BN254_MIN_BLOCKS) k_segment_sum(const void* pts, size_t groups, int len, int chunk, void* out) {
  cta_lockstep_set(false);  // no barriers in this kernel; the flag is read by the shared field routines
  int nch = (len + chunk - 1) / chunk;
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = t < groups * (size_t)nch;
  size_t g = live ? t / nch : 0;
  int c = live ? (int)(t % nch) : 0;
  int first = c * chunk, cnt = live ? min(chunk, len - first) : 0;
  J acc; f_set_zero(acc.x); f_set_zero(acc.y); f_set_zero(acc.z);
  for (int j = 0; j < cnt; j++) {
    A e; load_struct(e, pts, g * (size_t)len + first + j);
    if (!aff_is_inf(e)) jac_add_aff(acc, acc, e);
  }
  A r;
  jac_to_aff_inv<InvCta>(r, acc);  // one inversion per CTA
  if (live) store_struct(out, t, r);
}
// out[i] = -in[i]
template <typename A>
__global__ void k_neg_points(const void* in, size_t n, void* out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  A p; load_struct(p, in, i);
  p.y = f_neg(p.y);
  store_struct(out, i, p);
}


}  // namespace

namespace launch {

#define BY_GROUP(g, call1, call2) do { if ((g) == 1) { call1; } else { call2; } } while (0)
size_t subset_sum_max_bytes() { return 200 * 1024; }
cudaError_t group_init() {
  cudaError_t e = cudaFuncSetAttribute(k_subset_sum<G1Jac, G1Aff>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)subset_sum_max_bytes());
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_subset_sum<G2Jac, G2Aff>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)subset_sum_max_bytes());
  return e;
}
#define BY_GROUP(g, call1, call2) do { if ((g) == 1) { call1; } else { call2; } } while (0)
void aff_add(int g, const void* a, const void* b, size_t n, void* out, cudaStream_t s) {
  BY_GROUP(g, (BN_LAUNCH, k_aff_add<G1Jac, G1Aff><<<grid_for(n), kBlock, 0, s>>>(a, b, n, out)),
           (BN_LAUNCH, k_aff_add<G2Jac, G2Aff><<<grid_for(n), kBlock, 0, s>>>(a, b, n, out)));
}
void subset_sum(int g, const void* U, int m, const uint8_t* sel, size_t n, void* out, cudaStream_t s) {
  BY_GROUP(g, (BN_LAUNCH, k_subset_sum<G1Jac, G1Aff><<<grid_for(n), kBlock, (size_t)(m + 1) * sizeof(G1Aff), s>>>(static_cast<const G1Aff*>(U), m, sel, n, out)),
           (BN_LAUNCH, k_subset_sum<G2Jac, G2Aff><<<grid_for(n), kBlock, (size_t)(m + 1) * sizeof(G2Aff), s>>>(static_cast<const G2Aff*>(U), m, sel, n, out)));
}
void subset_sum_tab(int g, const void* U, int m, const uint8_t* sel, size_t n, void* out, void* table, cudaStream_t s) {
  int tt = ((m + 7) / 8) * 256;
  BY_GROUP(g, (BN_LAUNCH, k_subset_table<G1Jac, G1Aff><<<grid_for(tt), kBlock, 0, s>>>(static_cast<const G1Aff*>(U), m, static_cast<G1Aff*>(table))),
           (BN_LAUNCH, k_subset_table<G2Jac, G2Aff><<<grid_for(tt), kBlock, 0, s>>>(static_cast<const G2Aff*>(U), m, static_cast<G2Aff*>(table))));
  BY_GROUP(g, (BN_LAUNCH, k_subset_sum_tab<G1Jac, G1Aff><<<grid_for(n), kBlock, 0, s>>>(static_cast<const G1Aff*>(U), static_cast<const G1Aff*>(table), m, sel, n, out)),
           (BN_LAUNCH, k_subset_sum_tab<G2Jac, G2Aff><<<grid_for(n), kBlock, 0, s>>>(static_cast<const G2Aff*>(U), static_cast<const G2Aff*>(table), m, sel, n, out)));
}
void segment_sum(int g, const void* pts, size_t groups, int len, int chunk, void* out, cudaStream_t s) {
  size_t threads = groups * (size_t)((len + chunk - 1) / chunk);
  BY_GROUP(g, (BN_LAUNCH, k_segment_sum<G1Jac, G1Aff><<<grid_for(threads), kBlock, 0, s>>>(pts, groups, len, chunk, out)),
           (BN_LAUNCH, k_segment_sum<G2Jac, G2Aff><<<grid_for(threads), kBlock, 0, s>>>(pts, groups, len, chunk, out)));
}
void neg_points(int g, const void* in, size_t n, void* out, cudaStream_t s) {
  BY_GROUP(g, (BN_LAUNCH, k_neg_points<G1Aff><<<grid_for(n), kBlock, 0, s>>>(in, n, out)), (BN_LAUNCH, k_neg_points<G2Aff><<<grid_for(n), kBlock, 0, s>>>(in, n, out)));
}

}  // namespace launch
}  // namespace bn254
