// INT32 multiply pipe microbenchmark for B200 (sm_100a): the measured roofline denominator for the
// BN254 engine (SURVEY.md §8d asks for measured mad.lo/mad.hi/mad.wide rates, not the theoretical
// 148 x 64 x clk).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o imad_peak imad_peak.cu
// Run on the GPU box: ./imad_peak > gpurun_out/imad_peak.json
#include <cstdio>
#include <cstdint>
#include <string>
#include <cuda_runtime.h>
#include "../../gopairingbasedcryptography_b200/csrc/fp.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int ITERS = 4096;
constexpr int CH = 8;  // independent chains per thread

// each kernel: ITERS iterations x CH independent ops of the class under test
__global__ void k_imad_lo(uint32_t* out, uint32_t a, uint32_t b) {
  uint32_t x[CH];
  for (int i = 0; i < CH; i++) x[i] = threadIdx.x + i;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < CH; i++) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
  }
  uint32_t s = 0; for (int i = 0; i < CH; i++) s ^= x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_imad_hi(uint32_t* out, uint32_t a, uint32_t b) {
  uint32_t x[CH];
  for (int i = 0; i < CH; i++) x[i] = threadIdx.x + i;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < CH; i++) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
  }
  uint32_t s = 0; for (int i = 0; i < CH; i++) s ^= x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_imad_wide(uint32_t* out, uint32_t a, uint32_t b) {
  uint64_t x[CH];
  for (int i = 0; i < CH; i++) x[i] = threadIdx.x + i;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < CH; i++) { uint32_t hi = (uint32_t)(x[(i + 1) % CH] >> 32); asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x[i]) : "r"(hi), "r"(b)); }
  }
  uint64_t s = 0; for (int i = 0; i < CH; i++) s ^= x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)(s ^ (s >> 32));
}
// carry-chained wide MACs exactly as the Montgomery reduce row emits them (4 WIDE per chain)
__global__ void k_imad_wide_x(uint32_t* out, uint32_t m0) {
  using namespace bn254;
  uint32_t e[8], o[8];
  for (int i = 0; i < 8; i++) { e[i] = threadIdx.x + i; o[i] = threadIdx.x * 3 + i; }
  uint32_t m = m0 + threadIdx.x;
  for (int it = 0; it < ITERS / 2; it++) {
    o[0] = mad_lo_cc(P1, m, o[0]);  o[1] = madc_hi_cc(P1, m, o[1]);
    o[2] = madc_lo_cc(P3, m, o[2]); o[3] = madc_hi_cc(P3, m, o[3]);
    o[4] = madc_lo_cc(P5, m, o[4]); o[5] = madc_hi_cc(P5, m, o[5]);
    o[6] = madc_lo_cc(P7, m, o[6]); o[7] = madc_hi(P7, m, o[7]);
    e[0] = mad_lo_cc(P0, m, e[0]);  e[1] = madc_hi_cc(P0, m, e[1]);
    e[2] = madc_lo_cc(P2, m, e[2]); e[3] = madc_hi_cc(P2, m, e[3]);
    e[4] = madc_lo_cc(P4, m, e[4]); e[5] = madc_hi_cc(P4, m, e[5]);
    e[6] = madc_lo_cc(P6, m, e[6]); e[7] = madc_hi(P6, m, e[7]);
  }
  uint32_t s = 0; for (int i = 0; i < 8; i++) s ^= e[i] ^ o[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_iadd3(uint32_t* out, uint32_t a, uint32_t b) {
  uint32_t x[CH];
  for (int i = 0; i < CH; i++) x[i] = threadIdx.x + i;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < CH; i++) asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(x[(i + 1) % CH]));
  }
  uint32_t s = 0; for (int i = 0; i < CH; i++) s ^= x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + b + a;
}
// 1:1 mix: does the ALU pipe co-issue under a saturated IMAD pipe?
__global__ void k_mix(uint32_t* out, uint32_t a, uint32_t b) {
  uint32_t x[CH], y[CH];
  for (int i = 0; i < CH; i++) { x[i] = threadIdx.x + i; y[i] = threadIdx.x - i; }
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < CH; i++) {
      asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
      asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(y[i]) : "r"(y[(i + 1) % CH]));
    }
  }
  uint32_t s = 0; for (int i = 0; i < CH; i++) s ^= x[i] ^ y[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// dependent chain of Montgomery products: Fp-mul throughput at a given occupancy
template <int ILP, int V>
__global__ void k_fp_mul(uint32_t* out, int iters) {
  using namespace bn254;
  Fp x[ILP], y;
  for (int k = 0; k < ILP; k++) for (int i = 0; i < 8; i++) x[k].l[i] = (threadIdx.x * 2654435761u + i * 40503u + k) & 0x0fffffffu;
  for (int i = 0; i < 8; i++) y.l[i] = (blockIdx.x * 97u + i * 7919u + 1) & 0x0fffffffu;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int k = 0; k < ILP; k++) x[k] = fp_mul_v<V>(x[k], y);
  }
  uint32_t s = 0; for (int k = 0; k < ILP; k++) for (int i = 0; i < 8; i++) s ^= x[k].l[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float time_kernel(F launch) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; i++) launch();
  cudaEventRecord(e0);
  const int reps = 10;
  for (int i = 0; i < reps; i++) launch();
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms / reps;
}

int main(int argc, char** argv) {
  bool quick = argc > 1 && std::string(argv[1]) == "--quick";
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  uint32_t* out; CK(cudaMalloc(&out, sizeof(uint32_t) * sms * 16 * 1024));
  const int threads = 256, blocks = sms * 8;
  double total_threads = (double)threads * blocks;
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_rate_khz\": %d,\n", prop.name, sms, clk_khz);
  if (quick) {  // only the carry-chained IMAD.WIDE kernel: the live roofline denominator for bench.py
    float ms = time_kernel([&] { k_imad_wide_x<<<blocks, threads>>>(out, 7); });
    double ops = total_threads * (ITERS / 2) * 8;
    printf(" \"imad_wide_x_chain\": {\"ms\": %.4f, \"ops_per_s\": %.4e, \"ops_per_clk_per_sm_at_max_clock\": %.2f}}\n", ms,
           ops / (ms * 1e-3), ops / (ms * 1e-3) / sms / (clk_khz * 1e3));
    return 0;
  }
  struct { const char* name; float ms; double ops; } r[6];
  r[0] = {"imad_lo", time_kernel([&] { k_imad_lo<<<blocks, threads>>>(out, 3, 5); }), total_threads * ITERS * CH};
  r[1] = {"imad_hi", time_kernel([&] { k_imad_hi<<<blocks, threads>>>(out, 3, 5); }), total_threads * ITERS * CH};
  r[2] = {"imad_wide", time_kernel([&] { k_imad_wide<<<blocks, threads>>>(out, 3, 5); }), total_threads * ITERS * CH};
  r[3] = {"imad_wide_x_chain", time_kernel([&] { k_imad_wide_x<<<blocks, threads>>>(out, 7); }), total_threads * (ITERS / 2) * 8};
  r[4] = {"iadd", time_kernel([&] { k_iadd3<<<blocks, threads>>>(out, 3, 5); }), total_threads * ITERS * CH};
  r[5] = {"mix_imad_plus_iadd(imad count)", time_kernel([&] { k_mix<<<blocks, threads>>>(out, 3, 5); }), total_threads * ITERS * CH};
  CK(cudaGetLastError());
  for (int i = 0; i < 6; i++)
    printf(" \"%s\": {\"ms\": %.4f, \"ops_per_s\": %.4e, \"ops_per_clk_per_sm_at_max_clock\": %.2f},\n", r[i].name, r[i].ms,
           r[i].ops / (r[i].ms * 1e-3), r[i].ops / (r[i].ms * 1e-3) / sms / (clk_khz * 1e3));
  // Fp-mul throughput vs occupancy
  printf(" \"fp_mul\": [\n");
  int cfgs[][3] = {{128, 1, 1}, {128, 2, 1}, {128, 4, 1}, {256, 4, 1}, {256, 8, 1}, {128, 1, 2}, {128, 2, 2}, {128, 4, 2}, {128, 1, 4}, {128, 2, 4}, {128, 4, 4}};
  int ncfg = sizeof(cfgs) / sizeof(cfgs[0]);
  for (int v = 0; v < 2; v++)
  for (int c = 0; c < ncfg; c++) {
    int th = cfgs[c][0], bps = cfgs[c][1], ilp = cfgs[c][2];
    int iters = 2048;
    int nb = sms * bps;
    float ms;
    if (v == 0) {
      if (ilp == 1) ms = time_kernel([&] { k_fp_mul<1, 0><<<nb, th>>>(out, iters); });
      else if (ilp == 2) ms = time_kernel([&] { k_fp_mul<2, 0><<<nb, th>>>(out, iters); });
      else ms = time_kernel([&] { k_fp_mul<4, 0><<<nb, th>>>(out, iters); });
    } else {
      if (ilp == 1) ms = time_kernel([&] { k_fp_mul<1, 1><<<nb, th>>>(out, iters); });
      else if (ilp == 2) ms = time_kernel([&] { k_fp_mul<2, 1><<<nb, th>>>(out, iters); });
      else ms = time_kernel([&] { k_fp_mul<4, 1><<<nb, th>>>(out, iters); });
    }
    double muls = (double)nb * th * iters * ilp;
    printf("  {\"variant\": %d, \"threads\": %d, \"blocks_per_sm\": %d, \"ilp\": %d, \"ms\": %.4f, \"fp_mul_per_s\": %.4e}%s\n", v, th, bps, ilp, ms, muls / (ms * 1e-3), (v == 1 && c + 1 == ncfg) ? "" : ",");
  }
  printf(" ]}\n");
  CK(cudaDeviceSynchronize());
  return 0;
}
