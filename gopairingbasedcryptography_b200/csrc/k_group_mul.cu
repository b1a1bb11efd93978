// Group family, variable-base scalar multiplication: G1 / G2 2-dimensional GLV ladder, G2 4-dimensional GLS ladder.
// One translation unit per kernel group: adding the GLS kernel to the common unit moved k_fixed_mul<G2>'s code (45 -> 37 M/s).
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

}  // namespace

namespace launch {

#define BY_GROUP(g, call1, call2) do { if ((g) == 1) { call1; } else { call2; } } while (0)
size_t g2_gls_scratch_bytes(size_t n) { return n * (size_t)kGlsSliceFp2 * sizeof(Fp2); }
void scalar_mul_g2_gls(const void* base, size_t base_stride, const void* scalars, size_t n, void* out, void* scratch, cudaStream_t s) {
  BN_LAUNCH, k_scalar_mul_g2_gls<<<grid_for(n), kBlock, 0, s>>>(base, base_stride, scalars, n, out, static_cast<Fp2*>(scratch));
}
void scalar_mul(int g, const void* base, size_t base_stride, const void* scalars, size_t n, void* out, cudaStream_t s) {
  BY_GROUP(g, (BN_LAUNCH, k_scalar_mul<G1Jac, G1Aff><<<grid_for(n), kBlock, 0, s>>>(base, base_stride, scalars, n, out)),
           (BN_LAUNCH, k_scalar_mul<G2Jac, G2Aff><<<grid_for(n), kBlock, 0, s>>>(base, base_stride, scalars, n, out)));
}

}  // namespace launch
}  // namespace bn254
