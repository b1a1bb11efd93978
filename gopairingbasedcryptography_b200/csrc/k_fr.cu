// Scalar-field (Fr) feeders of the batch entry points -- SURVEY.md §8a row 11.
//
// The reference prepares the scalars it hands to ScalarMultiplication / GT.Exp with gnark's fr.Element on the host:
//   * bibe/afp25_bibe/afp25_bibe_utils.go:14-43 computePolynomialCoeffs: f(X) = prod (X - id_i), once per Digest and --
//     with one root removed -- again per Decrypt (O(B^2) Fr products each time);
//   * utils/compute_lagrange_basis.go:8-30 ComputeLagrangeBasis: one Fr inversion per factor, |S|^2 - |S| inversions
//     per threshold gate (access/tree/access_tree_node.go:155, fibe/sw05_fibe_common.go:316).
// Here: f is expanded once per identity batch on the GPU; the quotient f(X) / (X - id) of every decryption is an O(B)
// synthetic division (one warp per identity, the coefficients written straight into the MSM kernel's scalar layout,
// so they never cross PCIe); the Lagrange basis of a whole set costs ONE inversion (Montgomery's trick).
//
// fr.Element layout = gnark's: 4 x u64 little-endian limbs, Montgomery form R = 2^256 mod r, fully reduced.
#include "kcommon.cuh"

#include <vector>

namespace bn254 {
namespace {

struct Fr { uint32_t l[8]; };
// r, R mod r, R^2 mod r, -r^-1 mod 2^32 (python: see tests/test_fr_feeders.py, which re-derives them)
#define FR_CONST __host__ __device__ static inline
FR_CONST uint32_t fr_mod(int i) {
  const uint32_t m[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
  return m[i];
}
FR_CONST Fr fr_one() { return Fr{{0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u, 0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u}}; }
constexpr uint32_t kFrInv32 = 0xefffffffu;

FR_CONST Fr fr_zero() { Fr z; for (int i = 0; i < 8; i++) z.l[i] = 0; return z; }
FR_CONST bool fr_eq(const Fr& a, const Fr& b) { uint32_t o = 0; for (int i = 0; i < 8; i++) o |= a.l[i] ^ b.l[i]; return o == 0; }
FR_CONST bool fr_geq_mod(const uint32_t* t) {
  for (int i = 7; i >= 0; i--) { uint32_t m = fr_mod(i); if (t[i] != m) return t[i] > m; }
  return true;
}
FR_CONST Fr fr_add(const Fr& a, const Fr& b) {
  Fr z; uint64_t c = 0;
  for (int i = 0; i < 8; i++) { c += (uint64_t)a.l[i] + b.l[i]; z.l[i] = (uint32_t)c; c >>= 32; }
  if (fr_geq_mod(z.l)) { uint64_t bw = 0; for (int i = 0; i < 8; i++) { uint64_t d = (uint64_t)z.l[i] - fr_mod(i) - bw; z.l[i] = (uint32_t)d; bw = (d >> 32) & 1; } }
  return z;  // a + b < 2r < 2^255: no carry out of 8 limbs
}
FR_CONST Fr fr_sub(const Fr& a, const Fr& b) {
  Fr z; uint64_t bw = 0;
  for (int i = 0; i < 8; i++) { uint64_t d = (uint64_t)a.l[i] - b.l[i] - bw; z.l[i] = (uint32_t)d; bw = (d >> 32) & 1; }
  if (bw) { uint64_t c = 0; for (int i = 0; i < 8; i++) { c += (uint64_t)z.l[i] + fr_mod(i); z.l[i] = (uint32_t)c; c >>= 32; } }
  return z;
}
FR_CONST Fr fr_neg(const Fr& a) { return fr_sub(fr_zero(), a); }
// Montgomery product a*b/R mod r (CIOS, 32-bit digits, 64-bit accumulation); canonical result
FR_CONST Fr fr_mul(const Fr& a, const Fr& b) {
  uint32_t t[10];
  for (int i = 0; i < 10; i++) t[i] = 0;
  for (int i = 0; i < 8; i++) {
    uint64_t c = 0;
    for (int j = 0; j < 8; j++) { c += (uint64_t)a.l[j] * b.l[i] + t[j]; t[j] = (uint32_t)c; c >>= 32; }
    c += t[8]; t[8] = (uint32_t)c; t[9] = (uint32_t)(c >> 32);
    uint32_t m = t[0] * kFrInv32;
    c = ((uint64_t)m * fr_mod(0) + t[0]) >> 32;
    for (int j = 1; j < 8; j++) { c += (uint64_t)m * fr_mod(j) + t[j]; t[j - 1] = (uint32_t)c; c >>= 32; }
    c += t[8]; t[7] = (uint32_t)c; t[8] = t[9] + (uint32_t)(c >> 32);
  }
  Fr z;
  if (t[8] || fr_geq_mod(t)) { uint64_t bw = 0; for (int i = 0; i < 8; i++) { uint64_t d = (uint64_t)t[i] - fr_mod(i) - bw; z.l[i] = (uint32_t)d; bw = (d >> 32) & 1; } }
  else for (int i = 0; i < 8; i++) z.l[i] = t[i];
  return z;
}
// Montgomery -> regular form (the big.Int value, little-endian): multiply by 1
FR_CONST Fr fr_to_regular(const Fr& a) { Fr one = fr_zero(); one.l[0] = 1; return fr_mul(a, one); }
// a^(r-2); Inverse(0) = 0 as in gnark
static Fr fr_inv_host(const Fr& a) {
  uint32_t e[8];
  for (int i = 0; i < 8; i++) e[i] = fr_mod(i);
  e[0] -= 2;  // r - 2 (no borrow: the low limb is 0xf0000001)
  Fr acc = fr_one(), b = a;
  for (int i = 0; i < 254; i++) {
    if ((e[i >> 5] >> (i & 31)) & 1u) acc = fr_mul(acc, b);
    b = fr_mul(b, b);
  }
  return acc;
}

__device__ __forceinline__ Fr fr_load(const void* base, size_t i) {
  Fr v;
  const uint4* p = reinterpret_cast<const uint4*>(static_cast<const char*>(base) + i * 32);
  uint4 lo = __ldg(p), hi = __ldg(p + 1);
  v.l[0] = lo.x; v.l[1] = lo.y; v.l[2] = lo.z; v.l[3] = lo.w; v.l[4] = hi.x; v.l[5] = hi.y; v.l[6] = hi.z; v.l[7] = hi.w;
  return v;
}
__device__ __forceinline__ void fr_store(void* base, size_t i, const Fr& v) {
  uint4* p = reinterpret_cast<uint4*>(static_cast<char*>(base) + i * 32);
  p[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
  p[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
__device__ __forceinline__ Fr fr_shfl(const Fr& v, int src) {
  Fr r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.l[i] = __shfl_sync(0xffffffffu, v.l[i], src);
  return r;
}

// f(X) = prod_{i < n} (X - root_i): coefficients c_0 .. c_n (Montgomery), one CTA, the running polynomial in shared
// memory (double-buffered), one round per root: c'_t = c_{t-1} - root * c_t.
__global__ void __launch_bounds__(1024) k_fr_poly_from_roots(const void* roots, int n, void* coeffs) {
  extern __shared__ uint4 fr_smem[];
  Fr* buf[2] = {reinterpret_cast<Fr*>(fr_smem), reinterpret_cast<Fr*>(fr_smem) + (n + 1)};
  for (int t = threadIdx.x; t <= n; t += blockDim.x) buf[0][t] = t == 0 ? fr_one() : fr_zero();
  __syncthreads();
  int cur = 0;
  for (int i = 0; i < n; i++) {
    Fr root = fr_load(roots, i);
    const Fr* src = buf[cur];
    Fr* dst = buf[cur ^ 1];
    for (int t = threadIdx.x; t <= i + 1; t += blockDim.x) {
      Fr lo = t > 0 ? src[t - 1] : fr_zero();
      Fr v = t <= i ? fr_sub(lo, fr_mul(root, src[t])) : lo;
      dst[t] = v;
    }
    __syncthreads();
    cur ^= 1;
  }
  for (int t = threadIdx.x; t <= n; t += blockDim.x) fr_store(coeffs, t, buf[cur][t]);
}

// q(X) = f(X) / (X - id) for every id of a batch, f of degree n with f(id) = 0 (the division is exact; the remainder
// is not checked -- the reference does not check membership by value either, only by list position).
// Synthetic division Q[k-1] = f[k] + id * Q[k], Q[n] = 0, one WARP per id: lane s owns the segment Q[s*seg .. s*seg+seg),
//   pass 1: the segment with carry-in 0 -> A_s = local Q[s*seg];
//   carries: C_s = A_s + id^seg * C_{s+1} walked down the lanes;
//   pass 2: the segment again with its true carry-in, written in REGULAR form (big.Int value, little-endian) at
//           out[v * n + k]: the scalar layout of bn254_msm_batch (coefficient k multiplies point k = [tau^k]1).
__global__ void __launch_bounds__(128) k_fr_quotient_coeffs(const void* f, int n, const void* ids, size_t nvec, void* out) {
  size_t v = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (v >= nvec) return;  // whole warps leave together
  int lane = threadIdx.x & 31;
  int seg = (n + 31) / 32;
  Fr id = fr_load(ids, v);
  int lo = lane * seg, hi = min(lo + seg, n);  // Q indices [lo, hi)
  // pass 1
  Fr q = fr_zero();
  for (int k = hi; k > lo; k--) q = fr_add(fr_load(f, k), fr_mul(id, q));  // Q[k-1]
  // id^seg
  Fr pw = fr_one(), b = id;
  for (int e = seg; e; e >>= 1) { if (e & 1) pw = fr_mul(pw, b); b = fr_mul(b, b); }
  // carries, from the top lane down: every lane computes the same chain, lane s keeps C_{s+1}
  Fr carry_in = fr_zero(), c = fr_zero();
  for (int s = 31; s >= 0; s--) {
    Fr a_s = fr_shfl(q, s);
    if (lane == s) carry_in = c;
    // lanes whose segment is empty or short (n not a multiple of 32): A_s already accounts for its own length via
    // id^(hi-lo); the chain below assumes full segments, so short segments are handled by the per-lane power
    int len_s = min(s * seg + seg, n) - min(s * seg, n);
    Fr p = pw;
    if (len_s != seg) { p = fr_one(); Fr bb = id; for (int e = len_s; e; e >>= 1) { if (e & 1) p = fr_mul(p, bb); bb = fr_mul(bb, bb); } }
    c = fr_add(a_s, fr_mul(p, c));
  }
  // pass 2
  q = carry_in;
  for (int k = hi; k > lo; k--) {
    q = fr_add(fr_load(f, k), fr_mul(id, q));
    fr_store(out, v * (size_t)n + (k - 1), fr_to_regular(q));
  }
}
// Montgomery fr.Element -> regular-form little-endian scalars (x.BigInt(new(big.Int)) of the reference's call sites)
__global__ void k_fr_to_scalars(const void* in, size_t n, void* out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  fr_store(out, i, fr_to_regular(fr_load(in, i)));
}

}  // namespace

namespace launch {

cudaError_t fr_poly_from_roots(const void* roots, size_t n, void* coeffs, cudaStream_t s) {
  size_t smem = 2 * (n + 1) * sizeof(Fr);
  if (n == 0 || smem > 200 * 1024) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(k_fr_poly_from_roots, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  BN_LAUNCH, k_fr_poly_from_roots<<<1, 1024, smem, s>>>(roots, (int)n, coeffs);
  return cudaSuccess;
}
void fr_quotient_coeffs(const void* f, size_t n, const void* ids, size_t nvec, void* out, cudaStream_t s) {
  BN_LAUNCH, k_fr_quotient_coeffs<<<(unsigned)((nvec * 32 + 127) / 128), 128, 0, s>>>(f, (int)n, ids, nvec, out);
}
void fr_to_scalars(const void* in, size_t n, void* out, cudaStream_t s) { BN_LAUNCH, k_fr_to_scalars<<<grid_for(n), kBlock, 0, s>>>(in, n, out); }

// ---- host side: Lagrange basis of a whole set with one inversion ------------------------------------------------
// out[i] = Delta_{s_i, S}(x) = prod_{j: s_j != s_i} (x - s_j) / (s_i - s_j)   (utils/compute_lagrange_basis.go:8-30:
// the reference skips equal VALUES, not equal positions; so does this)
void fr_lagrange_basis_host(const void* s_in, size_t n, const void* x_in, void* out) {
  const Fr* s = static_cast<const Fr*>(s_in);
  Fr x = *static_cast<const Fr*>(x_in);
  std::vector<Fr> num(n), den(n), pre(n);
  for (size_t i = 0; i < n; i++) {
    Fr nu = fr_one(), de = fr_one();
    for (size_t j = 0; j < n; j++) {
      if (fr_eq(s[i], s[j])) continue;
      nu = fr_mul(nu, fr_sub(x, s[j]));
      de = fr_mul(de, fr_sub(s[i], s[j]));
    }
    num[i] = nu; den[i] = de;
  }
  Fr run = fr_one();
  for (size_t i = 0; i < n; i++) { pre[i] = run; run = fr_mul(run, den[i]); }  // den[i] != 0: distinct values only
  Fr inv = fr_inv_host(run);
  Fr* o = static_cast<Fr*>(out);
  for (size_t i = n; i-- > 0;) {
    Fr di = fr_mul(inv, pre[i]);
    inv = fr_mul(inv, den[i]);
    o[i] = fr_mul(num[i], di);
  }
}
void fr_to_scalars_host(const void* in, size_t n, void* out) {
  const Fr* a = static_cast<const Fr*>(in);
  Fr* o = static_cast<Fr*>(out);
  for (size_t i = 0; i < n; i++) o[i] = fr_to_regular(a[i]);
}

}  // namespace launch
}  // namespace bn254
