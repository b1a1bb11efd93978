// Shared by the kernel translation units (k_*.cu): launch geometry and AoS element access.
// One translation unit per kernel family (pairing, group, GT, hash, lane-group VM): every family gets its own copy
// of the field / tower device functions, so code generation of one family cannot move when another family changes
// (round 1: adding the GT.Exp kernel to the single translation unit changed k_pair's Fp2 product body, -1.4 %).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#include "launch.h"

#ifndef BN254_BLOCK
#define BN254_BLOCK 128
#endif
#ifndef BN254_MIN_BLOCKS
#define BN254_MIN_BLOCKS 1
#endif

namespace bn254 {

constexpr int kBlock = BN254_BLOCK;
static_assert(kBlock == launch::kBlockThreads, "launch.h and the kernel build disagree on the CTA size");

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
// 32-byte little-endian scalar -> 8 x u32
__device__ __forceinline__ void load_scalar(uint32_t* s, const void* scalars, size_t i) {
  const uint4* sp = reinterpret_cast<const uint4*>(static_cast<const char*>(scalars) + i * 32);
  uint4 lo = __ldg(sp), hi = __ldg(sp + 1);
  s[0] = lo.x; s[1] = lo.y; s[2] = lo.z; s[3] = lo.w; s[4] = hi.x; s[5] = hi.y; s[6] = hi.z; s[7] = hi.w;
}
__device__ __forceinline__ bool cta_is_full(size_t n) { return ((size_t)blockIdx.x + 1) * blockDim.x <= n; }
inline unsigned grid_for(size_t n) { return (unsigned)((n + kBlock - 1) / kBlock); }

}  // namespace bn254
