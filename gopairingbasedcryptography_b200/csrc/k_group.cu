// Group-family kernels: G1 / G2 scalar multiplication (GLV, fixed base), affine addition, subset sums, segment sums,
// shared-point MSM.
#include "kcommon.cuh"
#include "curve.cuh"

namespace bn254 {
namespace {
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
// G2 variable base by the 4-dimensional GLS ladder (curve.cuh); scratch: one kGlsSliceFp2 x 64 B slice per thread
__global__ void __launch_bounds__(kBlock, BN254_MIN_BLOCKS) k_scalar_mul_g2_gls(const void* base, size_t base_stride, const void* scalars, size_t n, void* out, Fp2* scratch) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  G2Aff b;
  bool plain = i < n;
  if (plain) { load_struct(b, base, i * base_stride); plain = !aff_is_inf(b); }
  cta_lockstep_set(__syncthreads_and(plain) != 0);
  if (i >= n) return;
  uint32_t s[8];
  load_scalar(s, scalars, i);
  G2Aff r;
  scalar_mul_gls4(r, b, s, scratch + i * (size_t)kGlsSliceFp2);
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
  if (i >= n) return;
  int nbytes = (m + 7) / 8;
  const uint8_t* bits = sel + i * (size_t)nbytes;
  J acc;
  A u0 = U[0];
  if (aff_is_inf(u0)) { f_set_zero(acc.x); f_set_zero(acc.y); f_set_zero(acc.z); }
  else { acc.x = u0.x; acc.y = u0.y; f_set_one(acc.z); }
  for (int b = 0; b < nbytes; b++) {
    int v = bits[b];
    if (v) {
      A e; load_struct(e, table, (size_t)b * 256 + v);
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
// ---- shared-point MSM (AFP25 / GWWW25: many coefficient vectors over the SAME tau-power points) ------------------
// bibe/afp25_bibe/afp25_bibe_utils.go:45-55 computes sum_j [c_j] T_j as len independent ScalarMultiplications plus len
// affine Adds (one inversion each), once per ciphertext, always over the public tau-power points.  Here the points get
// per-point window tables once -- tables[(j * 32 + w) * 255 + d - 1] = [d * 2^(8w)] P_j, affine -- and every term of
// every vector is 32 mixed additions with no doubling: 352 Fp-mul per term instead of ~2 300 for a GLV ladder.
//
// Table build: one thread per (point, window).  B = [2^(8w)] P by doublings, normalised once; the 255 multiples by
// repeated mixed addition, kept Jacobian (X, Y in the table slot, Z in a scratch row) and normalised together with
// ONE inversion per thread (Montgomery's trick over the 255 Z values, prefix products in a second scratch row).
template <typename T> struct field_of;
template <> struct field_of<G1Aff> { typedef Fp type; };
template <> struct field_of<G2Aff> { typedef Fp2 type; };
constexpr int kMsmWindows = launch::kMsmWindows;
static_assert(kMsmWindows == kFixedWindows && launch::kMsmWindowBits == 8, "byte windows");

template <typename J, typename A>
__global__ void __launch_bounds__(kBlock, BN254_MIN_BLOCKS) k_msm_tables(const void* pts, size_t len, A* tables,
                                                                         typename field_of<A>::type* zs, typename field_of<A>::type* pf) {
  typedef typename field_of<A>::type F;
  cta_lockstep_set(false);
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= len * (size_t)kMsmWindows) return;
  size_t j = t / kMsmWindows;
  int w = (int)(t % kMsmWindows);
  A p; load_struct(p, pts, j);
  A* row = tables + t * kFixedEntries;
  F* zrow = zs + t * kFixedEntries;
  F* prow = pf + t * kFixedEntries;
  A zero; f_set_zero(zero.x); f_set_zero(zero.y);
  if (aff_is_inf(p)) { for (int d = 0; d < kFixedEntries; d++) row[d] = zero; return; }
  J b; b.x = p.x; b.y = p.y; f_set_one(b.z);
  for (int i = 0; i < 8 * w; i++) jac_dbl(b, b);
  A ba; jac_to_aff(ba, b);
  if (aff_is_inf(ba)) { for (int d = 0; d < kFixedEntries; d++) row[d] = zero; return; }  // only off the prime-order subgroup
  J acc; acc.x = ba.x; acc.y = ba.y; f_set_one(acc.z);
  F run; f_set_one(run);
  for (int d = 1; d <= kFixedEntries; d++) {
    if (d > 1) jac_add_aff(acc, acc, ba);
    bool inf = jac_is_inf(acc);  // d * B = 0: only for points of small order (never in G1 or the G2 subgroup)
    A e; e.x = acc.x; e.y = acc.y;
    F z = acc.z;
    if (inf) { e = zero; f_set_one(z); }
    row[d - 1] = e; zrow[d - 1] = z;
    run = f_mul(run, z);
    prow[d - 1] = run;
  }
  F inv = f_inv(run);
  for (int d = kFixedEntries; d >= 1; d--) {
    F zi = inv;
    if (d > 1) { F pr = prow[d - 2]; zi = f_mul(inv, pr); }
    F z = zrow[d - 1];
    inv = f_mul(inv, z);
    F zi2 = f_sqr(zi);
    A e = row[d - 1];
    e.x = f_mul(e.x, zi2);
    e.y = f_mul(e.y, f_mul(zi2, zi));
    row[d - 1] = e;  // (0, 0) stays (0, 0)
  }
}
// partial[v * nchunks + c] = sum over the chunk's points j of sum_w tables[j][w][byte_w(s[v][j])] (Jacobian)
template <typename J, typename A>
__global__ void __launch_bounds__(kBlock, BN254_MIN_BLOCKS) k_msm_partial(const A* __restrict__ tables, const void* scalars, size_t nvec, size_t len, int chunk, J* partial) {
  cta_lockstep_set(false);
  size_t nch = (len + chunk - 1) / chunk;
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nvec * nch) return;
  size_t v = t / nch, c = t % nch;
  size_t first = c * (size_t)chunk, last = first + chunk < len ? first + chunk : len;
  J acc; f_set_zero(acc.x); f_set_zero(acc.y); f_set_zero(acc.z);
  for (size_t j = first; j < last; j++) {
    uint32_t s[8];
    load_scalar(s, scalars, v * len + j);
    const A* tj = tables + j * (size_t)kMsmWindows * kFixedEntries;
    for (int w = 0; w < kMsmWindows; w++) {
      int d = (int)((s[w >> 2] >> ((w & 3) * 8)) & 0xFFu);
      if (d) {
        A e; load_struct(e, tj, (size_t)w * kFixedEntries + d - 1);
        if (!aff_is_inf(e)) jac_add_aff(acc, acc, e);
      }
    }
  }
  partial[t] = acc;
}
// out[g * nch + c] = sum of the c-th chunk of group g's `len` Jacobian points; the last pass (nch == 1) may write affine
template <typename J, typename A>
__global__ void __launch_bounds__(kBlock, BN254_MIN_BLOCKS) k_jac_sum(const J* in, size_t groups, int len, int chunk, J* out_j, void* out_a) {
  cta_lockstep_set(false);
  int nch = (len + chunk - 1) / chunk;
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= groups * (size_t)nch) return;
  size_t g = t / nch;
  int c = (int)(t % nch);
  int first = c * chunk, cnt = min(chunk, len - first);
  J acc = in[g * (size_t)len + first];
  for (int j = 1; j < cnt; j++) { J e = in[g * (size_t)len + first + j]; jac_add(acc, acc, e); }
  if (out_a) { A r; jac_to_aff(r, acc); store_struct(out_a, t, r); }
  else out_j[t] = acc;
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

static_assert(bn254::kFixedWindows == launch::kFixedWindows && bn254::kFixedEntries == launch::kFixedEntries, "launch.h out of date");
size_t subset_sum_max_bytes() { return 200 * 1024; }
cudaError_t group_init() {
  cudaError_t e = cudaFuncSetAttribute(k_subset_sum<G1Jac, G1Aff>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)subset_sum_max_bytes());
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_subset_sum<G2Jac, G2Aff>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)subset_sum_max_bytes());
  return e;
}
#define BY_GROUP(g, call1, call2) do { if ((g) == 1) { call1; } else { call2; } } while (0)
size_t g2_gls_scratch_bytes(size_t n) { return n * (size_t)kGlsSliceFp2 * sizeof(Fp2); }
void scalar_mul_g2_gls(const void* base, size_t base_stride, const void* scalars, size_t n, void* out, void* scratch, cudaStream_t s) {
  BN_LAUNCH, k_scalar_mul_g2_gls<<<grid_for(n), kBlock, 0, s>>>(base, base_stride, scalars, n, out, static_cast<Fp2*>(scratch));
}
void scalar_mul(int g, const void* base, size_t base_stride, const void* scalars, size_t n, void* out, cudaStream_t s) {
  BY_GROUP(g, (BN_LAUNCH, k_scalar_mul<G1Jac, G1Aff><<<grid_for(n), kBlock, 0, s>>>(base, base_stride, scalars, n, out)),
           (BN_LAUNCH, k_scalar_mul<G2Jac, G2Aff><<<grid_for(n), kBlock, 0, s>>>(base, base_stride, scalars, n, out)));
}
void fixed_mul(int g, const void* table, const void* scalars, size_t n, void* out, cudaStream_t s) {
  BY_GROUP(g, (BN_LAUNCH, k_fixed_mul<G1Jac, G1Aff><<<grid_for(n), kBlock, 0, s>>>(static_cast<const G1Aff*>(table), scalars, n, out)),
           (BN_LAUNCH, k_fixed_mul<G2Jac, G2Aff><<<grid_for(n), kBlock, 0, s>>>(static_cast<const G2Aff*>(table), scalars, n, out)));
}
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
void msm_tables(int g, const void* pts, size_t len, void* tables, void* zs, void* pf, cudaStream_t s) {
  size_t threads = len * (size_t)kMsmWindows;
  BY_GROUP(g, (BN_LAUNCH, k_msm_tables<G1Jac, G1Aff><<<grid_for(threads), kBlock, 0, s>>>(pts, len, static_cast<G1Aff*>(tables), static_cast<Fp*>(zs), static_cast<Fp*>(pf))),
           (BN_LAUNCH, k_msm_tables<G2Jac, G2Aff><<<grid_for(threads), kBlock, 0, s>>>(pts, len, static_cast<G2Aff*>(tables), static_cast<Fp2*>(zs), static_cast<Fp2*>(pf))));
}
void msm_partial(int g, const void* tables, const void* scalars, size_t nvec, size_t len, int chunk, void* partial, cudaStream_t s) {
  size_t threads = nvec * ((len + chunk - 1) / chunk);
  BY_GROUP(g, (BN_LAUNCH, k_msm_partial<G1Jac, G1Aff><<<grid_for(threads), kBlock, 0, s>>>(static_cast<const G1Aff*>(tables), scalars, nvec, len, chunk, static_cast<G1Jac*>(partial))),
           (BN_LAUNCH, k_msm_partial<G2Jac, G2Aff><<<grid_for(threads), kBlock, 0, s>>>(static_cast<const G2Aff*>(tables), scalars, nvec, len, chunk, static_cast<G2Jac*>(partial))));
}
void jac_sum(int g, const void* in, size_t groups, int len, int chunk, void* out_j, void* out_a, cudaStream_t s) {
  size_t threads = groups * (size_t)((len + chunk - 1) / chunk);
  BY_GROUP(g, (BN_LAUNCH, k_jac_sum<G1Jac, G1Aff><<<grid_for(threads), kBlock, 0, s>>>(static_cast<const G1Jac*>(in), groups, len, chunk, static_cast<G1Jac*>(out_j), out_a)),
           (BN_LAUNCH, k_jac_sum<G2Jac, G2Aff><<<grid_for(threads), kBlock, 0, s>>>(static_cast<const G2Jac*>(in), groups, len, chunk, static_cast<G2Jac*>(out_j), out_a)));
}
void neg_points(int g, const void* in, size_t n, void* out, cudaStream_t s) {
  BY_GROUP(g, (BN_LAUNCH, k_neg_points<G1Aff><<<grid_for(n), kBlock, 0, s>>>(in, n, out)), (BN_LAUNCH, k_neg_points<G2Aff><<<grid_for(n), kBlock, 0, s>>>(in, n, out)));
}

}  // namespace launch
}  // namespace bn254
