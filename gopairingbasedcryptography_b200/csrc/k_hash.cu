// Hash-to-curve kernels (gnark bn254.HashToG1 / HashToG2).
#include "kcommon.cuh"
#include "hash_to_curve.cuh"

namespace bn254 {
namespace {
// up to this many messages HashToG2 runs two lanes per message (about one warp per scheduler: latency-bound either way)
constexpr size_t kHashPairsMax = 8192;
// hash-to-curve: one message per thread (SHA-256 expand_message_xmd, SVDW map x2, add, G2 cofactor clearing)
template <int G>
__global__ void __launch_bounds__(kBlock, BN254_MIN_BLOCKS) k_hash_to_curve(const uint8_t* msgs, const uint64_t* off, size_t n, const uint8_t* dst,
                                                                           uint32_t dst_len, void* out) {
  cta_lockstep_set(false);  // the tower's lockstep barriers are off; the shared inversions below bring their own
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  // G1: the two inversions of a hash (SVDW, normalisation) are 40 % of its field products, and the CTA shares them
  // (InvCta: one Fermat ladder per 128 threads, a product tree in shared memory) -- 27.7 -> 31.8 M hashes/s -- so EVERY
  // thread walks the whole routine: threads past the end hash message 0 (n >= 1) and store nothing.  G2 keeps its own
  // ladders: there the inversions are 15 % of the work and the barriers between warps that drift apart (Jacobi rounds,
  // message lengths) cost what the sharing saves (9.86 vs 9.77 M/s measured).
  const bool live = i < n;
  if (G == 2 && !live) return;
  const size_t ii = live ? i : 0;
  const uint8_t* m = msgs + off[ii];
  size_t len = (size_t)(off[ii + 1] - off[ii]);
  if (G == 1) { G1Aff r; hash_to_g1<InvCta>(r, m, len, dst, dst_len); if (live) store_struct(out, i, r); }
  else { G2Aff r; hash_to_g2<InvThread>(r, m, len, dst, dst_len); store_struct(out, i, r); }
}

// Small batches are latency-bound (one thread needs ~4 ms per HashToG2 whatever the batch size), and the two SVDW
// maps of a message -- 57 % of its ~8 000 dependent Fp products -- are independent: TWO adjacent lanes share a message,
// each maps one of the two field elements, the odd lane hands its point to the even lane by warp shuffles, and the even
// lane adds, clears the cofactor and normalises.  Same arithmetic per message, bit-identical output.
__global__ void __launch_bounds__(kBlock, BN254_MIN_BLOCKS) k_hash_to_g2_pairs(const uint8_t* msgs, const uint64_t* off, size_t n, const uint8_t* dst,
                                                                              uint32_t dst_len, void* out) {
  cta_lockstep_set(false);
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t i = t >> 1;
  const int half = (int)(t & 1);
  const bool live = i < n;
  if (!live) i = 0;  // idle lane pairs hash message 0 so that every lane reaches the shuffles (n >= 1)
  const uint8_t* m = msgs + off[i];
  size_t len = (size_t)(off[i + 1] - off[i]);
  Fp u[4];
  hash_to_field<4>(u, m, len, dst, dst_len);
  Fp2 uu; uu.a0 = u[2 * half]; uu.a1 = u[2 * half + 1];
  G2Aff q, q1;
  map_to_curve_g2(q, uu);
  const uint32_t* src = reinterpret_cast<const uint32_t*>(&q);
  uint32_t* dstw = reinterpret_cast<uint32_t*>(&q1);
#pragma unroll
  for (int w = 0; w < (int)(sizeof(G2Aff) / 4); w++) dstw[w] = __shfl_down_sync(0xffffffffu, src[w], 1);
  if (half || !live) return;
  G2Jac s; s.x = q.x; s.y = q.y; s.z = fp2_one();
  jac_add_aff(s, s, q1);
  g2_clear_cofactor(s, s);
  G2Aff r;
  jac_to_aff(r, s);
  store_struct(out, i, r);
}

}  // namespace

namespace launch {

cudaError_t hash_init() { return cudaSuccess; }  // no tower scratch: launched without dynamic shared memory (G1: 90 registers, 5 CTAs per SM)
void hash_to_curve(int g, const uint8_t* msgs, const uint64_t* off, size_t n, const uint8_t* dst, uint32_t dst_len, void* out, cudaStream_t s) {
  if (g == 1) BN_LAUNCH, k_hash_to_curve<1><<<grid_for(n), kBlock, 0, s>>>(msgs, off, n, dst, dst_len, out);
  else if (n > 0 && n <= kHashPairsMax) BN_LAUNCH, k_hash_to_g2_pairs<<<grid_for(2 * n), kBlock, 0, s>>>(msgs, off, n, dst, dst_len, out);
  else BN_LAUNCH, k_hash_to_curve<2><<<grid_for(n), kBlock, 0, s>>>(msgs, off, n, dst, dst_len, out);
}

}  // namespace launch
}  // namespace bn254
