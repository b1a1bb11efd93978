// Can Tensor Memory serve as a per-thread scratchpad for big-integer tower values?  (DESIGN.md §7, round-2 option:
// a fourth CTA per SM needs 128 B/thread more fast scratch than shared memory has.)
// Each CTA of 128 threads allocates 128 TMEM columns (512 B per lane); every thread writes 64-byte "slots" to its own
// lane with tcgen05.st.32x32b.x16 and reads them back with tcgen05.ld.32x32b.x16.  Reports data integrity and the
// store->load round-trip time in cycles, for 1..4 CTAs per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_scratch tmem_scratch.cu && ./tmem_scratch
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kCols = 128;

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n"
      :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
         "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr) : "memory");
}

__global__ void __launch_bounds__(128) k_tmem(uint32_t* bad, unsigned long long* cycles, int iters) {
  __shared__ uint32_t base_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "l"((uint64_t)__cvta_generic_to_shared(&base_s)), "n"(kCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  const uint32_t base = base_s;
  const uint32_t lane_base = base + ((uint32_t)(warp * 32) << 16);  // warp w owns lanes 32w .. 32w+31
  uint32_t v[16], r[16];
  for (int i = 0; i < 16; i++) v[i] = threadIdx.x * 1000003u + blockIdx.x * 7919u + i;
  uint32_t errors = 0;
  // integrity: fill all 8 slots, read them back
  for (int s = 0; s < kCols / 16; s++) {
    for (int i = 0; i < 16; i++) v[i] += 0x01010101u * (s + 1);
    tmem_st16(lane_base + s * 16, v);
  }
  asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
  for (int i = 0; i < 16; i++) v[i] = threadIdx.x * 1000003u + blockIdx.x * 7919u + i;
  for (int s = 0; s < kCols / 16; s++) {
    for (int i = 0; i < 16; i++) v[i] += 0x01010101u * (s + 1);
    tmem_ld16(lane_base + s * 16, r);
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
    for (int i = 0; i < 16; i++) errors += (r[i] != v[i]);
  }
  // round trip: store one slot, wait, load it, wait, dependent update
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    tmem_st16(lane_base + (it & 7) * 16, v);
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
    tmem_ld16(lane_base + (it & 7) * 16, r);
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
    for (int i = 0; i < 16; i++) v[i] = r[i] + 1;
  }
  long long t1 = clock64();
  for (int i = 0; i < 16; i++) errors += (v[i] == 0xdeadbeefu);
  if (errors) atomicAdd(bad, errors);
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = (unsigned long long)(t1 - t0);
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(base), "n"(kCols));
}

int main() {
  uint32_t* bad; unsigned long long* cyc;
  cudaMalloc(&bad, 4); cudaMalloc(&cyc, 8);
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int iters = 2000;
  for (int per_sm = 1; per_sm <= 4; per_sm++) {
    cudaMemset(bad, 0, 4); cudaMemset(cyc, 0, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k_tmem<<<sms * per_sm, 128>>>(bad, cyc, iters);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    uint32_t hb = 0; unsigned long long hc = 0;
    cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost); cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
    printf("{\"ctas_per_sm\": %d, \"status\": \"%s\", \"mismatches\": %u, \"round_trip_cycles\": %.1f, \"kernel_ms\": %.3f}\n", per_sm,
           cudaGetErrorString(err), hb, (double)hc / iters, ms);
    if (err != cudaSuccess) return 1;
  }
  return 0;
}
