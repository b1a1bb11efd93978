// GT-family kernels: GT.Exp (generic, GT-proper GLV ladder, fixed base), GT.Mul / Div, raw Fp product.
#include "kcommon.cuh"
#include "curve.cuh"

namespace bn254 {
namespace {
#ifdef BN254_SMEM_SCRATCH
constexpr size_t kTowerSmem = (size_t)kBlock * kScratchStride;
#else
constexpr size_t kTowerSmem = 0;
#endif
using launch::kGtCycloTable;  // Fp12 entries of per-thread table space gt_cyclo_exp needs
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
// mode 0: a*b ; mode 1: a/b (generic Fp12 inverse) ; mode 2: a * conj(b) = a/b for b in the cyclotomic subgroup
template <int MODE>
__global__ void __launch_bounds__(kBlock, BN254_MIN_BLOCKS) k_gt_mul(const void* a, size_t a_stride, const void* b, size_t b_stride, size_t n, void* out) {
  cta_lockstep_set(false);
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fp12 x, y; load_struct(x, a, i * a_stride); load_struct(y, b, i * b_stride);
  if (MODE == 1) fp12_inv(y, y);
  if (MODE == 2) fp12_conj(y, y);
  fp12_mul(x, x, y);
  store_struct(out, i, x);
}
__global__ void __launch_bounds__(kBlock, BN254_MIN_BLOCKS) k_fp_mul(const void* a, const void* b, size_t n, void* out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fp x, y; load_struct(x, a, i); load_struct(y, b, i);
  x = fp_mul(x, y);
  store_struct(out, i, x);
}

}  // namespace

namespace launch {

cudaError_t gt_init() {
#ifdef BN254_SMEM_SCRATCH
  const void* kernels[] = {(const void*)k_gt_fixed_exp, (const void*)k_gt_exp<0>, (const void*)k_gt_exp<1>, (const void*)k_gt_mul<0>, (const void*)k_gt_mul<1>,
                           (const void*)k_gt_mul<2>};
  for (const void* k : kernels) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTowerSmem);
    if (e != cudaSuccess) return e;
  }
#endif
  return cudaSuccess;
}
int gt_wave_threads(int sms) { return sms * BN254_MIN_BLOCKS * kBlock; }
void gt_exp(int cyclo, const void* x, size_t x_stride, const void* k, size_t n, void* out, void* tabmem, cudaStream_t s) {
  if (cyclo) BN_LAUNCH, k_gt_exp<1><<<grid_for(n), kBlock, kTowerSmem, s>>>(x, x_stride, k, n, out, static_cast<Fp12*>(tabmem));
  else BN_LAUNCH, k_gt_exp<0><<<grid_for(n), kBlock, kTowerSmem, s>>>(x, x_stride, k, n, out, static_cast<Fp12*>(tabmem));
}
void gt_fixed_exp(const void* table, const void* k, size_t n, void* out, cudaStream_t s) {
  BN_LAUNCH, k_gt_fixed_exp<<<grid_for(n), kBlock, kTowerSmem, s>>>(static_cast<const Fp12*>(table), k, n, out);
}
void gt_mul(int mode, const void* a, size_t a_stride, const void* b, size_t b_stride, size_t n, void* out, cudaStream_t s) {
  if (mode == 0) BN_LAUNCH, k_gt_mul<0><<<grid_for(n), kBlock, kTowerSmem, s>>>(a, a_stride, b, b_stride, n, out);
  else if (mode == 1) BN_LAUNCH, k_gt_mul<1><<<grid_for(n), kBlock, kTowerSmem, s>>>(a, a_stride, b, b_stride, n, out);
  else BN_LAUNCH, k_gt_mul<2><<<grid_for(n), kBlock, kTowerSmem, s>>>(a, a_stride, b, b_stride, n, out);
}
void fp_mul(const void* a, const void* b, size_t n, void* out, cudaStream_t s) { BN_LAUNCH, k_fp_mul<<<grid_for(n), kBlock, 0, s>>>(a, b, n, out); }

}  // namespace launch
}  // namespace bn254
