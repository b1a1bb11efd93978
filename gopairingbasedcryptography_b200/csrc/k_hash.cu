// Hash-to-curve kernels (gnark bn254.HashToG1 / HashToG2).
#include "kcommon.cuh"
#include "hash_to_curve.cuh"

namespace bn254 {
namespace {
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

}  // namespace

namespace launch {

cudaError_t hash_init() { return cudaSuccess; }  // no tower scratch: launched without dynamic shared memory (G1: 90 registers, 5 CTAs per SM)
void hash_to_curve(int g, const uint8_t* msgs, const uint64_t* off, size_t n, const uint8_t* dst, uint32_t dst_len, void* out, cudaStream_t s) {
  if (g == 1) BN_LAUNCH, k_hash_to_curve<1><<<grid_for(n), kBlock, 0, s>>>(msgs, off, n, dst, dst_len, out);
  else BN_LAUNCH, k_hash_to_curve<2><<<grid_for(n), kBlock, 0, s>>>(msgs, off, n, dst, dst_len, out);
}

}  // namespace launch
}  // namespace bn254
