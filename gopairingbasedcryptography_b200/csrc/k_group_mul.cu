// Group family, variable-base scalar multiplication: G1 / G2 2-dimensional GLV ladder, G2 4-dimensional GLS ladder.
// One translation unit per kernel group: adding the GLS kernel to the common unit moved k_fixed_mul<G2>'s code (45 -> 37 M/s).
#include "kcommon.cuh"
#include "curve.cuh"

namespace bn254 {
namespace {
// Every thread of the CTA runs the SAME ladder -- threads past the end of the batch and infinity bases multiply the
// generator by their scalar and drop the result -- so the CTA shares its two inversions (curve.cuh: InvCta) and always
// runs in lockstep.
template <typename A> __device__ __forceinline__ A group_generator();
template <> __device__ __forceinline__ G1Aff group_generator<G1Aff>() { G1Aff g; g.x = G1_GEN_X; g.y = G1_GEN_Y; return g; }
template <> __device__ __forceinline__ G2Aff group_generator<G2Aff>() { G2Aff g; g.x = G2_GEN_X; g.y = G2_GEN_Y; return g; }

template <typename J, typename A>
__global__ void __launch_bounds__(kBlock, BN254_MIN_BLOCKS) k_scalar_mul(const void* base, size_t base_stride, const void* scalars, size_t n, void* out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < n;
  A b;
  uint32_t s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  bool inf = false;
  if (live) { load_struct(b, base, i * base_stride); inf = aff_is_inf(b); load_scalar(s, scalars, i); }
  if (!live || inf) b = group_generator<A>();
  cta_lockstep_set(true);
  A r;
  Fp beta = (sizeof(A) == sizeof(G1Aff)) ? GLV_BETA : GLV_BETA_G2;
  scalar_mul_glv<J, A, InvCta>(r, b, s, beta);
  if (!live) return;
  if (inf) { f_set_zero(r.x); f_set_zero(r.y); }
  store_struct(out, i, r);
}
// G2 variable base by the 4-dimensional GLS ladder (curve.cuh); scratch: one kGlsSliceFp2 x 64 B slice per THREAD of the grid
__global__ void __launch_bounds__(kBlock, BN254_MIN_BLOCKS) k_scalar_mul_g2_gls(const void* base, size_t base_stride, const void* scalars, size_t n, void* out, Fp2* scratch) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < n;
  G2Aff b;
  uint32_t s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  bool inf = false;
  if (live) { load_struct(b, base, i * base_stride); inf = aff_is_inf(b); load_scalar(s, scalars, i); }
  if (!live || inf) b = group_generator<G2Aff>();
  cta_lockstep_set(true);
  G2Aff r;
  // (per-thread inversions here: sharing them across the CTA measured 12.3 against 12.5 M mults/s -- at 3 CTAs per SM the
  // three warps waiting for the inverting one are missed more than the saved ladders)
  scalar_mul_gls4<InvThread>(r, b, s, scratch + i * (size_t)kGlsSliceFp2);
  if (!live) return;
  if (inf) { f_set_zero(r.x); f_set_zero(r.y); }
  store_struct(out, i, r);
}

}  // namespace

namespace launch {

#define BY_GROUP(g, call1, call2) do { if ((g) == 1) { call1; } else { call2; } } while (0)
size_t g2_gls_scratch_bytes(size_t n) { return n * (size_t)kGlsSliceFp2 * sizeof(Fp2); }  // callers add kBlockThreads items: idle threads of the last CTA run the ladder too
void scalar_mul_g2_gls(const void* base, size_t base_stride, const void* scalars, size_t n, void* out, void* scratch, cudaStream_t s) {
  BN_LAUNCH, k_scalar_mul_g2_gls<<<grid_for(n), kBlock, 0, s>>>(base, base_stride, scalars, n, out, static_cast<Fp2*>(scratch));
}
void scalar_mul(int g, const void* base, size_t base_stride, const void* scalars, size_t n, void* out, cudaStream_t s) {
  BY_GROUP(g, (BN_LAUNCH, k_scalar_mul<G1Jac, G1Aff><<<grid_for(n), kBlock, 0, s>>>(base, base_stride, scalars, n, out)),
           (BN_LAUNCH, k_scalar_mul<G2Jac, G2Aff><<<grid_for(n), kBlock, 0, s>>>(base, base_stride, scalars, n, out)));
}

}  // namespace launch
}  // namespace bn254
