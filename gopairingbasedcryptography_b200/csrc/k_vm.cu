// Lane-group ("tower VM") kernels: K lanes per pairing, state in shared memory.  See vm.cuh.
#include "kcommon.cuh"
#include "vm.cuh"

namespace bn254 {
namespace {
#ifndef BN254_VM_K
#define BN254_VM_K 3
#endif
#ifndef BN254_VM_WARPS
#define BN254_VM_WARPS 4
#endif
constexpr int kVmK = BN254_VM_K;
constexpr int kVmWarps = BN254_VM_WARPS;
constexpr int kVmGroups = 32 / kVmK;            // pairings per warp
constexpr int kVmNP = kVmGroups * kVmWarps;     // pairings per CTA
constexpr int kVmStride = kVmNP | 1;            // odd stride: sub-lanes of a group hit different bank quads
constexpr int kVmColdSlots = 64;                // cold slots reserved per pairing in the global scratch

#define VM_CAT_(a, b, c) a##b##c
#define VM_CAT(a, b, c) VM_CAT_(a, b, c)
#define VM_SYM(name, suffix) VM_CAT(vm::name##_K, BN254_VM_K, suffix)
#if BN254_VM_K != 3
#error "only the K = 3 programs are kept in-tree: run `python vmgen.py <K>` and add the includes for another lane width"
#endif
#define VM_INC_PAIR "vm_prog_pair_k3.inc"
#define VM_INC_MILLER "vm_prog_miller_k3.inc"
#define VM_INC_FINALEXP "vm_prog_finalexp_k3.inc"
__device__ const uint64_t kProgPair[] = {
#include VM_INC_PAIR
};
__device__ const uint64_t kProgMiller[] = {
#include VM_INC_MILLER
};
__device__ const uint64_t kProgFinalExp[] = {
#include VM_INC_FINALEXP
};

struct VmProgPair { static constexpr int rounds = VM_SYM(PAIR, _ROUNDS), nslots = VM_SYM(PAIR, _NSLOTS), ncold = VM_SYM(PAIR, _NCOLD), nin = 3;
  __device__ static const uint64_t* prog() { return kProgPair; }
  __device__ static int in(int i) { return VM_SYM(PAIR, _IN)[i]; } __device__ static int out(int i) { return VM_SYM(PAIR, _OUT)[i]; } };
struct VmProgMiller { static constexpr int rounds = VM_SYM(MILLER, _ROUNDS), nslots = VM_SYM(MILLER, _NSLOTS), ncold = VM_SYM(MILLER, _NCOLD), nin = 3;
  __device__ static const uint64_t* prog() { return kProgMiller; }
  __device__ static int in(int i) { return VM_SYM(MILLER, _IN)[i]; } __device__ static int out(int i) { return VM_SYM(MILLER, _OUT)[i]; } };
struct VmProgFinalExp { static constexpr int rounds = VM_SYM(FINALEXP, _ROUNDS), nslots = VM_SYM(FINALEXP, _NSLOTS), ncold = VM_SYM(FINALEXP, _NCOLD), nin = 6;
  __device__ static const uint64_t* prog() { return kProgFinalExp; }
  __device__ static int in(int i) { return VM_SYM(FINALEXP, _IN)[i]; } __device__ static int out(int i) { return VM_SYM(FINALEXP, _OUT)[i]; } };

template <typename PROG> constexpr size_t vm_smem_bytes() { return (size_t)PROG::nslots * 4 * kVmStride * sizeof(uint4); }

// Persistent CTAs: each warp owns kVmGroups pairings at a time; warps never synchronise with each other.
// PROG::nin == 3: inputs are (P, Q.x, Q.y) from the G1/G2 arrays; PROG::nin == 6: the six Fp2 of a GT.
template <typename PROG>
__global__ void __launch_bounds__(32 * kVmWarps) k_vm(const void* in0, const void* in1, size_t n, void* out, uint4* cold) {
  extern __shared__ uint4 vm_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane / kVmK, j = lane % kVmK;
  const bool lane_ok = g < kVmGroups;
  const int pid = warp * kVmGroups + (lane_ok ? g : 0);
  vm::SlotFile f;
  f.hot = vm_smem; f.nslots = PROG::nslots; f.hot_stride = kVmStride; f.pid = pid;
  f.cold = cold; f.cold_stride = gridDim.x * kVmNP; f.gpid = blockIdx.x * kVmNP + pid;
  const unsigned gmask = lane_ok ? (((1u << kVmK) - 1u) << (g * kVmK)) : 0u;
  for (size_t base = (size_t)blockIdx.x * kVmNP; base < n; base += (size_t)gridDim.x * kVmNP) {
    const size_t idx = base + pid;
    const bool active = lane_ok && idx < n;
    // ---- prologue: operands -> slots; pairs containing the point at infinity are flagged ----
    unsigned nzP = 0, nzQ = 0;
    for (int i = 0; i < PROG::nin; i++) {
      bool mine = active && (i % kVmK) == j;
      uint32_t nz = 0;
      if (mine) {
        const char* src;
        if (PROG::nin == 3) src = (i == 0) ? static_cast<const char*>(in0) + idx * 64 : static_cast<const char*>(in1) + idx * 128 + (i - 1) * 64;
        else src = static_cast<const char*>(in0) + idx * 384 + i * 64;
        Fp2 v;
        uint4* d = reinterpret_cast<uint4*>(&v);
#pragma unroll
        for (int c = 0; c < 4; c++) { d[c] = __ldg(reinterpret_cast<const uint4*>(src) + c); nz |= d[c].x | d[c].y | d[c].z | d[c].w; }
        vm::st_slot(f, PROG::in(i), v);
      }
      unsigned b = __ballot_sync(0xffffffffu, nz != 0);
      if (i == 0) nzP = b & gmask; else nzQ |= b & gmask;
    }
    const bool skip = (PROG::nin == 3) && (nzP == 0 || nzQ == 0);
    __syncwarp();
    vm::run<kVmK>(f, PROG::prog(), PROG::rounds, j, active);
    // ---- epilogue ----
    if (active) {
      for (int i = j; i < 6; i += kVmK) {
        Fp2 v;
        if (skip) { v = fp2_zero(); if (i == 0) v.a0 = fp_one(); }
        else v = vm::ld_slot(f, PROG::out(i));
        uint4* dst = reinterpret_cast<uint4*>(static_cast<char*>(out) + idx * 384 + i * 64);
        const uint4* sv = reinterpret_cast<const uint4*>(&v);
#pragma unroll
        for (int c = 0; c < 4; c++) dst[c] = sv[c];
      }
    }
    __syncwarp();
  }
}

static_assert(kVmNP == launch::kVmPairingsPerCta, "launch.h out of date");
static_assert(VmProgPair::ncold <= kVmColdSlots && VmProgMiller::ncold <= kVmColdSlots && VmProgFinalExp::ncold <= kVmColdSlots,
              "a generated program uses more cold slots than the global scratch reserves per pairing");

template <typename PROG>
cudaError_t vm_prepare_one(int* blocks) {
  cudaError_t e = cudaFuncSetAttribute(k_vm<PROG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)vm_smem_bytes<PROG>());
  if (e != cudaSuccess) return e;
  int nb = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_vm<PROG>, 32 * kVmWarps, vm_smem_bytes<PROG>());
  if (e != cudaSuccess) return e;
  if (nb < 1) return cudaErrorLaunchOutOfResources;
  *blocks = nb;
  return cudaSuccess;
}
template <typename PROG>
void vm_launch(const void* a, const void* b, size_t n, void* out, void* cold, int sms, int blocks_per_sm, cudaStream_t s) {
  size_t want = (n + kVmNP - 1) / kVmNP;
  unsigned grid = (unsigned)(want < (size_t)sms * blocks_per_sm ? want : (size_t)sms * blocks_per_sm);
  BN_LAUNCH, k_vm<PROG><<<grid, 32 * kVmWarps, vm_smem_bytes<PROG>(), s>>>(a, b, n, out, static_cast<uint4*>(cold));
}

}  // namespace

namespace launch {

cudaError_t vm_prepare(int* blocks_per_sm) {
  cudaError_t e = vm_prepare_one<VmProgPair>(&blocks_per_sm[kVmPair]);
  if (e == cudaSuccess) e = vm_prepare_one<VmProgMiller>(&blocks_per_sm[kVmMiller]);
  if (e == cudaSuccess) e = vm_prepare_one<VmProgFinalExp>(&blocks_per_sm[kVmFinalExp]);
  return e;
}
size_t vm_cold_bytes(int sms, const int* b) {
  int maxb = b[0] > b[1] ? b[0] : b[1];
  if (b[2] > maxb) maxb = b[2];
  return (size_t)kVmColdSlots * 4 * sizeof(uint4) * (size_t)sms * maxb * kVmNP;
}
void vm_run(int prog, const void* a, const void* b, size_t n, void* out, void* cold, int sms, const int* blocks_per_sm, cudaStream_t s) {
  if (prog == kVmPair) vm_launch<VmProgPair>(a, b, n, out, cold, sms, blocks_per_sm[kVmPair], s);
  else if (prog == kVmMiller) vm_launch<VmProgMiller>(a, b, n, out, cold, sms, blocks_per_sm[kVmMiller], s);
  else vm_launch<VmProgFinalExp>(a, b, n, out, cold, sms, blocks_per_sm[kVmFinalExp], s);
}

}  // namespace launch
}  // namespace bn254
