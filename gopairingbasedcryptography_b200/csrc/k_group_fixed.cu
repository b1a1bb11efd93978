// Group family, table-driven kernels: fixed-base window tables, shared-point MSM (tables, partial sums, Jacobian tree).
#include "kcommon.cuh"
#include "curve.cuh"

namespace bn254 {
namespace {
// fixed base: 32 windowed mixed additions from a precomputed affine table (L2-resident, 0.5-1 MB)
template <typename J, typename A>
__global__ void __launch_bounds__(kBlock, BN254_MIN_BLOCKS) k_fixed_mul(const A* table, const void* scalars, size_t n, void* out) {
  cta_lockstep_set(false);  // no lockstep barriers in this kernel; the flag is read by the shared field routines
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < n;
  uint32_t s[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // threads past the end add nothing and hand z = 0 to the CTA's shared inversion
  if (live) load_scalar(s, scalars, i);
  A r;
  scalar_mul_fixed<J, A, InvCta>(r, table, s);
  if (live) store_struct(out, i, r);
}
// Large batches: FOUR consecutive scalars per thread, one inversion for the four results (the final inversion is half of
// a G1 fixed-base multiplication: 380 of 732 Fp products).  Items [4t, 4t + 4) are contiguous for the thread.
constexpr int kFixedItems = 4;
constexpr size_t kFixedBatchMin = (size_t)1 << 16;
template <typename J, typename A>
__global__ void __launch_bounds__(kBlock, 4) k_fixed_mul_x4(const A* table, const void* scalars, size_t n, void* out) {
  cta_lockstep_set(false);
  size_t first = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * kFixedItems;
  if (first >= n) return;
  int count = (int)(n - first < (size_t)kFixedItems ? n - first : (size_t)kFixedItems);
  J acc[kFixedItems];
  for (int b = 0; b < count; b++) {
    uint32_t s[8];
    load_scalar(s, scalars, first + b);
    scalar_mul_fixed_jac<J, A>(acc[b], table, s);
  }
  A r[kFixedItems];
  jac_to_aff_batch<kFixedItems, J, A>(r, acc, count);
  for (int b = 0; b < count; b++) store_struct(out, first + b, r[b]);
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
// Jacobian tree over groups of `width` consecutive threads of a CTA (width a power of two <= kBlock) through shared
// memory: log2(width) dependent additions instead of width - 1.  The first thread of every group ends with the sum.
// Every thread of the CTA must call it (barriers); idle lanes pass the point at infinity.
template <typename J>
__device__ __forceinline__ void cta_jac_tree(J& acc, int width) {
  J* sm = reinterpret_cast<J*>(bn_dyn_smem);
  const int tid = threadIdx.x, j = tid & (width - 1);
  for (int step = 1; step < width; step <<= 1) {
    sm[tid] = acc;
    __syncthreads();
    if ((j & (2 * step - 1)) == 0) { J e = sm[tid + step]; jac_add(acc, acc, e); }
    __syncthreads();
  }
}
// partial[v * nchunks + c] = sum over the chunk's points j of sum_w tables[j][w][byte_w(s[v][j])] (Jacobian).
// TREE (the launch guarantees nchunks % kBlock == 0, so a CTA lies inside one vector and no thread is idle): the CTA's
// kBlock sums are added by a shared-memory tree and ONE partial per CTA is written -- partial[blockIdx.x].
template <typename J, typename A, bool TREE>
__global__ void __launch_bounds__(kBlock, BN254_MIN_BLOCKS) k_msm_partial(const A* __restrict__ tables, const void* scalars, size_t nvec, size_t len, int chunk, J* partial) {
  cta_lockstep_set(false);
  size_t nch = (len + chunk - 1) / chunk;
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (!TREE && t >= nvec * nch) return;
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
  if (TREE) {
    cta_jac_tree(acc, kBlock);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc;
  } else {
    partial[t] = acc;
  }
}
// out_a[g] = affine sum of group g's `len` Jacobian points, len a power of two <= kBlock: one thread per point, tree
// per group (3 dependent additions for len = 8 instead of 7), one inversion per group.
template <typename J, typename A>
__global__ void __launch_bounds__(kBlock, BN254_MIN_BLOCKS) k_jac_tree(const J* in, size_t groups, int len, void* out_a) {
  cta_lockstep_set(false);
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  J acc;
  if (t < groups * (size_t)len) acc = in[t];
  else { f_set_zero(acc.x); f_set_zero(acc.y); f_set_zero(acc.z); }
  cta_jac_tree(acc, len);
  if (t < groups * (size_t)len && (threadIdx.x & (len - 1)) == 0) { A r; jac_to_aff(r, acc); store_struct(out_a, t / len, r); }
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

}  // namespace

namespace launch {

static_assert(bn254::kFixedWindows == launch::kFixedWindows && bn254::kFixedEntries == launch::kFixedEntries, "launch.h out of date");
#define BY_GROUP(g, call1, call2) do { if ((g) == 1) { call1; } else { call2; } } while (0)
void fixed_mul(int g, const void* table, const void* scalars, size_t n, void* out, cudaStream_t s) {
  // G1 only: on G2 the inversion is a quarter of the work and four Fp2 Jacobian points per thread double the stack
  // (measured: G1 86 -> 104 M/s at 2^18, G2 45 -> 35 M/s at 2^17)
  if (g == 1 && n >= kFixedBatchMin) {  // enough items to keep every scheduler busy with four per thread
    size_t threads = (n + kFixedItems - 1) / kFixedItems;
    BN_LAUNCH, k_fixed_mul_x4<G1Jac, G1Aff><<<grid_for(threads), kBlock, 0, s>>>(static_cast<const G1Aff*>(table), scalars, n, out);
    return;
  }
  BY_GROUP(g, (BN_LAUNCH, k_fixed_mul<G1Jac, G1Aff><<<grid_for(n), kBlock, 0, s>>>(static_cast<const G1Aff*>(table), scalars, n, out)),
           (BN_LAUNCH, k_fixed_mul<G2Jac, G2Aff><<<grid_for(n), kBlock, 0, s>>>(static_cast<const G2Aff*>(table), scalars, n, out)));
}
void msm_tables(int g, const void* pts, size_t len, void* tables, void* zs, void* pf, cudaStream_t s) {
  size_t threads = len * (size_t)kMsmWindows;
  BY_GROUP(g, (BN_LAUNCH, k_msm_tables<G1Jac, G1Aff><<<grid_for(threads), kBlock, 0, s>>>(pts, len, static_cast<G1Aff*>(tables), static_cast<Fp*>(zs), static_cast<Fp*>(pf))),
           (BN_LAUNCH, k_msm_tables<G2Jac, G2Aff><<<grid_for(threads), kBlock, 0, s>>>(pts, len, static_cast<G2Aff*>(tables), static_cast<Fp2*>(zs), static_cast<Fp2*>(pf))));
}
void msm_partial(int g, const void* tables, const void* scalars, size_t nvec, size_t len, int chunk, void* partial, cudaStream_t s) {
  size_t threads = nvec * ((len + chunk - 1) / chunk);
  BY_GROUP(g, (BN_LAUNCH, k_msm_partial<G1Jac, G1Aff, false><<<grid_for(threads), kBlock, 0, s>>>(static_cast<const G1Aff*>(tables), scalars, nvec, len, chunk, static_cast<G1Jac*>(partial))),
           (BN_LAUNCH, k_msm_partial<G2Jac, G2Aff, false><<<grid_for(threads), kBlock, 0, s>>>(static_cast<const G2Aff*>(tables), scalars, nvec, len, chunk, static_cast<G2Jac*>(partial))));
}
// one partial per CTA (requires ((len + chunk - 1) / chunk) % kBlockThreads == 0)
void msm_partial_tree(int g, const void* tables, const void* scalars, size_t nvec, size_t len, int chunk, void* partial, cudaStream_t s) {
  size_t threads = nvec * ((len + chunk - 1) / chunk);
  BY_GROUP(g, (BN_LAUNCH, k_msm_partial<G1Jac, G1Aff, true><<<grid_for(threads), kBlock, kBlock * sizeof(G1Jac), s>>>(static_cast<const G1Aff*>(tables), scalars, nvec, len, chunk, static_cast<G1Jac*>(partial))),
           (BN_LAUNCH, k_msm_partial<G2Jac, G2Aff, true><<<grid_for(threads), kBlock, kBlock * sizeof(G2Jac), s>>>(static_cast<const G2Aff*>(tables), scalars, nvec, len, chunk, static_cast<G2Jac*>(partial))));
}
// out_a[g] = affine sum of `len` Jacobian points per group, len a power of two <= kBlockThreads
void jac_tree(int g, const void* in, size_t groups, int len, void* out_a, cudaStream_t s) {
  size_t threads = groups * (size_t)len;
  BY_GROUP(g, (BN_LAUNCH, k_jac_tree<G1Jac, G1Aff><<<grid_for(threads), kBlock, kBlock * sizeof(G1Jac), s>>>(static_cast<const G1Jac*>(in), groups, len, out_a)),
           (BN_LAUNCH, k_jac_tree<G2Jac, G2Aff><<<grid_for(threads), kBlock, kBlock * sizeof(G2Jac), s>>>(static_cast<const G2Jac*>(in), groups, len, out_a)));
}
void jac_sum(int g, const void* in, size_t groups, int len, int chunk, void* out_j, void* out_a, cudaStream_t s) {
  size_t threads = groups * (size_t)((len + chunk - 1) / chunk);
  BY_GROUP(g, (BN_LAUNCH, k_jac_sum<G1Jac, G1Aff><<<grid_for(threads), kBlock, 0, s>>>(static_cast<const G1Jac*>(in), groups, len, chunk, static_cast<G1Jac*>(out_j), out_a)),
           (BN_LAUNCH, k_jac_sum<G2Jac, G2Aff><<<grid_for(threads), kBlock, 0, s>>>(static_cast<const G2Jac*>(in), groups, len, chunk, static_cast<G2Jac*>(out_j), out_a)));
}

}  // namespace launch
}  // namespace bn254
