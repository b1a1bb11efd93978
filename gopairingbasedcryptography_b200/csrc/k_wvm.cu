// Warp-VM kernels: one WARP per pairing / Miller loop / final exponentiation -- the latency path.  See wvm.cuh and
// wvmgen.py.  A 1-element bn254.Pair (the reference's own call shape) finishes in well under a millisecond instead of
// ~6 ms on the K = 3 lane-group kernels and ~20 ms on one thread.
#include "kcommon.cuh"
#include "wvm.cuh"
#include "wvm_prog_meta.cuh"

namespace bn254 {
namespace {

__device__ const uint32_t kWvmMiller[] = {
#include "wvm_prog_miller.inc"
};
__device__ const uint32_t kWvmMiller2[] = {
#include "wvm_prog_miller2.inc"
};
__device__ const uint32_t kWvmFinalExp[] = {
#include "wvm_prog_finalexp.inc"
};
constexpr int kWvmWarps = 4;
constexpr int kWvmSlots12 = wvm::MILLER_NSLOTS > wvm::FINALEXP_NSLOTS ? wvm::MILLER_NSLOTS : wvm::FINALEXP_NSLOTS;
constexpr int kWvmSlots = kWvmSlots12 > wvm::MILLER2_NSLOTS ? kWvmSlots12 : wvm::MILLER2_NSLOTS;
constexpr size_t kWvmSmem = (size_t)kWvmWarps * kWvmSlots * sizeof(Fp);
static_assert(kWvmSlots <= 1024, "10-bit slot fields");

__device__ __forceinline__ void wvm_consts(Fp* slots, const uint16_t* cslot, const Fp* cval, int n, int lane) {
  for (int i = lane; i < n; i += 32) wvm::st_slot(slots, cslot[i], cval[i]);
}

// MODE 0: Miller loop (in0 = P, in1 = Q), 1: pairing, 2: final exponentiation (in0 = GT)
template <int MODE>
__global__ void __launch_bounds__(32 * kWvmWarps) k_wvm(const void* in0, const void* in1, size_t n, void* out) {
  extern __shared__ uint4 wvm_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  Fp* slots = reinterpret_cast<Fp*>(wvm_smem) + (size_t)warp * kWvmSlots;
  for (size_t idx = (size_t)blockIdx.x * kWvmWarps + warp; idx < n; idx += (size_t)gridDim.x * kWvmWarps) {
    // ---- prologue: constants and operands -> slots; a pair containing the point at infinity yields 1 ----
    Fp v = fp_zero();
    bool skip = false;
    if (MODE != 2) {
      if (lane < 2) v = fp_ld(*reinterpret_cast<const Fp*>(static_cast<const char*>(in0) + idx * 64 + lane * 32));
      else if (lane < 6) v = fp_ld(*reinterpret_cast<const Fp*>(static_cast<const char*>(in1) + idx * 128 + (lane - 2) * 32));
      unsigned nz = __ballot_sync(0xffffffffu, !fp_is_zero(v));
      skip = (nz & 0x3u) == 0 || (nz & 0x3Cu) == 0;
      wvm_consts(slots, wvm::MILLER_CONST_SLOT, wvm::MILLER_CONST_VAL, wvm::MILLER_NCONST, lane);
      if (lane < 6) wvm::st_slot(slots, wvm::MILLER_IN[lane], v);
    } else {
      if (lane < 12) v = fp_ld(*reinterpret_cast<const Fp*>(static_cast<const char*>(in0) + idx * 384 + lane * 32));
      wvm_consts(slots, wvm::FINALEXP_CONST_SLOT, wvm::FINALEXP_CONST_VAL, wvm::FINALEXP_NCONST, lane);
      if (lane < 12) wvm::st_slot(slots, wvm::FINALEXP_IN[lane], v);
    }
    __syncwarp();
    if (!skip) {
      if (MODE != 2) {
        wvm::run(slots, reinterpret_cast<const uint4*>(kWvmMiller), wvm::MILLER_ROUNDS, lane);
        if (lane < 12) v = wvm::ld_slot(slots, wvm::MILLER_OUT[lane]);
        __syncwarp();
      }
      if (MODE == 1) {  // hand the Miller value to the final-exponentiation program
        wvm_consts(slots, wvm::FINALEXP_CONST_SLOT, wvm::FINALEXP_CONST_VAL, wvm::FINALEXP_NCONST, lane);
        if (lane < 12) wvm::st_slot(slots, wvm::FINALEXP_IN[lane], v);
        __syncwarp();
      }
      if (MODE >= 1) {
        wvm::run(slots, reinterpret_cast<const uint4*>(kWvmFinalExp), wvm::FINALEXP_ROUNDS, lane);
        if (lane < 12) v = wvm::ld_slot(slots, wvm::FINALEXP_OUT[lane]);
        __syncwarp();
      }
    } else {
      v = lane == 0 ? fp_one() : fp_zero();
    }
    if (lane < 12) {
      uint4* dst = reinterpret_cast<uint4*>(static_cast<char*>(out) + idx * 384 + lane * 32);
      const uint4* sv = reinterpret_cast<const uint4*>(&v);
      dst[0] = sv[0]; dst[1] = sv[1];
    }
    __syncwarp();
  }
}

// Product of the Miller loops of TWO pairs per item with shared squarings (in0 = P[2n], in1 = Q[2n]; the BLS verification
// and every 2-pair PairingCheck): one warp instead of two, no separate product kernel.  The two-pair program cannot skip a
// pair, so an item with a point at infinity in one pair runs the single-pair program on the other (gnark skips such pairs).
__global__ void __launch_bounds__(32 * kWvmWarps) k_wvm_miller2(const void* in0, const void* in1, size_t n, void* out) {
  extern __shared__ uint4 wvm_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  Fp* slots = reinterpret_cast<Fp*>(wvm_smem) + (size_t)warp * kWvmSlots;
  for (size_t idx = (size_t)blockIdx.x * kWvmWarps + warp; idx < n; idx += (size_t)gridDim.x * kWvmWarps) {
    const char* p0 = static_cast<const char*>(in0) + idx * 128;
    const char* q0 = static_cast<const char*>(in1) + idx * 256;
    Fp v = fp_zero();
    if (lane < 12) {  // MILLER2_IN order: per pair (P.x, P.y, Q.x.a0, Q.x.a1, Q.y.a0, Q.y.a1)
      int j = lane / 6, k = lane % 6;
      v = k < 2 ? fp_ld(*reinterpret_cast<const Fp*>(p0 + j * 64 + k * 32)) : fp_ld(*reinterpret_cast<const Fp*>(q0 + j * 128 + (k - 2) * 32));
    }
    unsigned nz = __ballot_sync(0xffffffffu, !fp_is_zero(v));
    const bool skip0 = (nz & 0x3u) == 0 || (nz & 0x3Cu) == 0, skip1 = (nz & 0xC0u) == 0 || (nz & 0xF00u) == 0;
    if (!skip0 && !skip1) {
      wvm_consts(slots, wvm::MILLER2_CONST_SLOT, wvm::MILLER2_CONST_VAL, wvm::MILLER2_NCONST, lane);
      if (lane < 12) wvm::st_slot(slots, wvm::MILLER2_IN[lane], v);
      __syncwarp();
      wvm::run(slots, reinterpret_cast<const uint4*>(kWvmMiller2), wvm::MILLER2_ROUNDS, lane);
      if (lane < 12) v = wvm::ld_slot(slots, wvm::MILLER2_OUT[lane]);
    } else if (skip0 && skip1) {
      v = lane == 0 ? fp_one() : fp_zero();
    } else {
      const int j = skip0 ? 1 : 0;
      v = fp_zero();
      if (lane < 2) v = fp_ld(*reinterpret_cast<const Fp*>(p0 + j * 64 + lane * 32));
      else if (lane < 6) v = fp_ld(*reinterpret_cast<const Fp*>(q0 + j * 128 + (lane - 2) * 32));
      wvm_consts(slots, wvm::MILLER_CONST_SLOT, wvm::MILLER_CONST_VAL, wvm::MILLER_NCONST, lane);
      if (lane < 6) wvm::st_slot(slots, wvm::MILLER_IN[lane], v);
      __syncwarp();
      wvm::run(slots, reinterpret_cast<const uint4*>(kWvmMiller), wvm::MILLER_ROUNDS, lane);
      if (lane < 12) v = wvm::ld_slot(slots, wvm::MILLER_OUT[lane]);
    }
    __syncwarp();
    if (lane < 12) {
      uint4* dst = reinterpret_cast<uint4*>(static_cast<char*>(out) + idx * 384 + lane * 32);
      const uint4* sv = reinterpret_cast<const uint4*>(&v);
      dst[0] = sv[0]; dst[1] = sv[1];
    }
    __syncwarp();
  }
}

template <int MODE>
cudaError_t prepare_one(int* blocks) {
  cudaError_t e = cudaFuncSetAttribute(k_wvm<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWvmSmem);
  if (e != cudaSuccess) return e;
  int nb = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_wvm<MODE>, 32 * kWvmWarps, kWvmSmem);
  if (e != cudaSuccess) return e;
  if (nb < 1) return cudaErrorLaunchOutOfResources;
  *blocks = nb;
  return cudaSuccess;
}
template <int MODE>
void launch_one(const void* a, const void* b, size_t n, void* out, int sms, int blocks_per_sm, cudaStream_t s) {
  size_t want = (n + kWvmWarps - 1) / kWvmWarps;
  size_t cap = (size_t)sms * blocks_per_sm;
  BN_LAUNCH, k_wvm<MODE><<<(unsigned)(want < cap ? want : cap), 32 * kWvmWarps, kWvmSmem, s>>>(a, b, n, out);
}

}  // namespace

namespace launch {

cudaError_t wvm_prepare(int* blocks_per_sm) {
  cudaError_t e = prepare_one<0>(&blocks_per_sm[kVmMiller]);
  if (e == cudaSuccess) e = prepare_one<1>(&blocks_per_sm[kVmPair]);
  if (e == cudaSuccess) e = prepare_one<2>(&blocks_per_sm[kVmFinalExp]);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_wvm_miller2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWvmSmem);
  if (e == cudaSuccess) {
    int nb = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_wvm_miller2, 32 * kWvmWarps, kWvmSmem);
    if (e == cudaSuccess && nb < 1) e = cudaErrorLaunchOutOfResources;
    blocks_per_sm[kVmMiller2] = nb;
  }
  return e;
}
int wvm_items_per_cta() { return kWvmWarps; }
#ifdef BN254_WVM_PROFILE
extern "C" void bn254_wvm_profile(unsigned long long* out8) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out8, wvm::wvm_prof, sizeof(unsigned long long) * 8);
  unsigned long long z[8] = {0};
  cudaMemcpyToSymbol(wvm::wvm_prof, z, sizeof(z));
}
#endif
void wvm_run(int prog, const void* a, const void* b, size_t n, void* out, int sms, const int* blocks_per_sm, cudaStream_t s) {
  if (prog == kVmMiller2) {
    size_t want = (n + kWvmWarps - 1) / kWvmWarps, cap = (size_t)sms * blocks_per_sm[kVmMiller2];
    BN_LAUNCH, k_wvm_miller2<<<(unsigned)(want < cap ? want : cap), 32 * kWvmWarps, kWvmSmem, s>>>(a, b, n, out);
  } else if (prog == kVmMiller) launch_one<0>(a, b, n, out, sms, blocks_per_sm[kVmMiller], s);
  else if (prog == kVmPair) launch_one<1>(a, b, n, out, sms, blocks_per_sm[kVmPair], s);
  else launch_one<2>(a, b, n, out, sms, blocks_per_sm[kVmFinalExp], s);
}

}  // namespace launch
}  // namespace bn254
