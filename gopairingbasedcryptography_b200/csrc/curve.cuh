// G1 / G2 group arithmetic for BN254 (Jacobian coordinates inside, canonical affine outside).
// Replaces (reference side): gnark-crypto v0.19.0 ecc/bn254/{g1,g2}.go ScalarMultiplication /
// ScalarMultiplicationBase / Add / Neg, called at signature/bls01_signature/bls_signature.go:45,63,
// ibe/waters05_ibe/waters05_ibe.go:224-237, cpabe/bsw07/bsw07_cpabe.go:69-121,149-160,
// bibe/afp25_bibe/afp25_bibe_utils.go:45-55.  gnark returns canonical affine points, so any correct
// algorithm is bit-exact.
#pragma once
#include "pairing.cuh"

namespace bn254 {

struct G1Jac { Fp x, y, z; };
struct G2Jac { Fp2 x, y, z; };

// field-generic helpers (overloads on Fp / Fp2)
BN_HD Fp f_add(const Fp& a, const Fp& b) { return fp_add(a, b); }
BN_HD Fp f_sub(const Fp& a, const Fp& b) { return fp_sub(a, b); }
BN_HD Fp f_dbl(const Fp& a) { return fp_dbl(a); }
BN_HD Fp f_neg(const Fp& a) { return fp_neg(a); }
// Products and inversions go through BY-VALUE out-of-line bodies (operands and results in registers).  The earlier
// by-reference form (fp_mul_ool(z, a, b) with a temporary z) is avoided: nvcc 12.9 was seen to give the temporary
// the stack slot of a still-live operand (t = sqr(x); y = mul(t, x) read x^2 for x -- hash_to_curve.cuh, found by
// the GPU parity test against the oracle); values whose address is never taken cannot be hit by that.
// base^e for a CONSTANT 256-bit exponent (the same in every lane: no SIMT divergence) by a sliding window of four bits
// over the odd powers base^1 .. base^15: ~254 squarings + ~51 products + 8 for the table instead of the 254 + ~127 of
// the bit-by-bit ladder (p - 2, (p - 1)/2 and (p + 1)/4 are dense).  0^e = 0 for e > 0.  Same value, fewer products:
// the inversions of the group kernels and the fixed-exponent ladders of hash-to-curve (four per G1 map, five per G2 map).
BN_NOINLINE Fp fp_pow_win(Fp b, const uint32_t* e) {
  Fp tab[8];
  {
    Fp b2 = FP_MUL(b, b);
    tab[0] = b;
#pragma unroll
    for (int i = 1; i < 8; i++) tab[i] = FP_MUL(tab[i - 1], b2);
  }
  Fp acc = fp_one();
  bool started = false;
  int i = 255;
  while (i >= 0 && !((e[i >> 5] >> (i & 31)) & 1u)) i--;
  while (i >= 0) {
    if (!((e[i >> 5] >> (i & 31)) & 1u)) { acc = FP_MUL(acc, acc); i--; continue; }
    int j = i - 3 < 0 ? 0 : i - 3;
    while (!((e[j >> 5] >> (j & 31)) & 1u)) j++;  // the window [j, i] ends in a set bit
    int w = 0;
    for (int t = i; t >= j; t--) w = (w << 1) | (int)((e[t >> 5] >> (t & 31)) & 1u);
    if (started) {
      for (int t = i; t >= j; t--) acc = FP_MUL(acc, acc);
      acc = FP_MUL(acc, tab[w >> 1]);
    } else {
      acc = tab[w >> 1];
      started = true;
    }
    i = j - 1;
  }
  return acc;
}
BN_NOINLINE Fp fp_inv_bv(Fp a) { return fp_pow_win(a, FP_PM2); }  // a^(p-2); inv(0) = 0 like gnark's Inverse
BN_HD Fp f_mul(const Fp& a, const Fp& b) { return FP_MUL(a, b); }
BN_HD Fp f_sqr(const Fp& a) { return FP_MUL(a, a); }
BN_HD Fp f_inv(const Fp& a) { return fp_inv_bv(a); }
BN_HD bool f_is_zero(const Fp& a) { return fp_is_zero(a); }
BN_HD void f_set_one(Fp& a) { a = fp_one(); }
BN_HD void f_set_zero(Fp& a) { a = fp_zero(); }
BN_HD Fp2 f_add(const Fp2& a, const Fp2& b) { return fp2_add(a, b); }
BN_HD Fp2 f_sub(const Fp2& a, const Fp2& b) { return fp2_sub(a, b); }
BN_HD Fp2 f_dbl(const Fp2& a) { return fp2_dbl(a); }
BN_HD Fp2 f_neg(const Fp2& a) { return fp2_neg(a); }
BN_HD Fp2 f_mul(const Fp2& a, const Fp2& b) { return FP2_MUL(a, b); }
BN_HD Fp2 f_sqr(const Fp2& a) { return FP2_SQR(a); }
BN_HD Fp2 f_inv(const Fp2& a) {
  Fp n = fp_add(f_sqr(a.a0), f_sqr(a.a1));
  Fp ni = f_inv(n);
  Fp2 r; r.a0 = f_mul(a.a0, ni); r.a1 = fp_neg(f_mul(a.a1, ni));
  return r;
}
BN_HD bool f_is_zero(const Fp2& a) { return fp2_is_zero(a); }
BN_HD void f_set_one(Fp2& a) { a = fp2_one(); }
BN_HD void f_set_zero(Fp2& a) { a = fp2_zero(); }

// ---- who inverts ----------------------------------------------------------------------------------------------------
// InvThread: every thread runs its own Fermat ladder (380 dependent Fp products).
// InvCta (kernels only): the kBlock threads of a CTA share ONE ladder.  SIMT lanes run in lock step, so batching inside a
// warp saves nothing -- a ladder costs a warp the same time for 1 or 32 lanes -- but across the 4 warps of a CTA it does:
// a product tree over the 128 values in shared memory (7 levels up, one inversion by thread 0, 7 levels down with two
// products each) costs the CTA ~400 warp-level products instead of 4 x 380.  Every thread of the CTA must call it the same
// number of times (barriers); inv(0) = 0 as in f_inv.
#if defined(__CUDACC__)
#define BN_MEMBER __device__ __forceinline__ static
#else
#define BN_MEMBER static inline
#endif
struct InvThread {
  BN_MEMBER Fp inv(const Fp& a) { return f_inv(a); }
  BN_MEMBER Fp2 inv(const Fp2& a) { return f_inv(a); }
};
#if defined(__CUDACC__)
#ifndef BN254_BLOCK
#define BN254_BLOCK 128
#endif
BN_NOINLINE Fp cta_fp_inv(Fp z) {
  constexpr int B = BN254_BLOCK;
  __shared__ Fp tree[2 * B];  // level l (B >> l entries) starts at 2B - (2B >> l)
  const int tid = threadIdx.x;
  const bool zero = fp_is_zero(z);
  tree[tid] = zero ? fp_one() : z;
  __syncthreads();
  int lo = 0;
  for (int w = B >> 1; w >= 1; w >>= 1) {  // up: parent = left * right
    int up = lo + 2 * w;
    if (tid < w) tree[up + tid] = FP_MUL(tree[lo + 2 * tid], tree[lo + 2 * tid + 1]);
    __syncthreads();
    lo = up;
  }
  if (tid == 0) tree[lo] = fp_inv_bv(tree[lo]);
  __syncthreads();
  for (int w = 1; w <= B >> 1; w <<= 1) {  // down: inv(left) = inv(parent) * right, inv(right) = inv(parent) * left
    int dn = lo - 2 * w;
    if (tid < w) {
      Fp pinv = tree[lo + tid], l = tree[dn + 2 * tid], r = tree[dn + 2 * tid + 1];
      tree[dn + 2 * tid] = FP_MUL(pinv, r);
      tree[dn + 2 * tid + 1] = FP_MUL(pinv, l);
    }
    __syncthreads();
    lo = dn;
  }
  Fp r = tree[tid];
  __syncthreads();  // the tree is reused by the next call
  return zero ? fp_zero() : r;
}
struct InvCta {
  BN_MEMBER Fp inv(const Fp& a) { return cta_fp_inv(a); }
  BN_MEMBER Fp2 inv(const Fp2& a) {  // conj(a) / norm(a): the norm is the shared inversion
    Fp ni = cta_fp_inv(fp_add(f_sqr(a.a0), f_sqr(a.a1)));
    Fp2 r; r.a0 = f_mul(a.a0, ni); r.a1 = fp_neg(f_mul(a.a1, ni));
    return r;
  }
};
#endif

#ifdef BN254_OOL_JAC
#define BN_JAC BN_NOINLINE
#else
#define BN_JAC BN_HD
#endif
template <typename J>
BN_HD bool jac_is_inf(const J& p) { return f_is_zero(p.z); }

// a = 0 doubling (dbl-2009-l)
template <typename J>
BN_JAC void jac_dbl(J& r, const J& p) {
  if (jac_is_inf(p)) { r = p; return; }
  auto A = f_sqr(p.x), B = f_sqr(p.y), C = f_sqr(B);
  auto D = f_dbl(f_sub(f_sub(f_sqr(f_add(p.x, B)), A), C));
  auto E = f_add(f_dbl(A), A);
  auto F = f_sqr(E);
  auto z3 = f_dbl(f_mul(p.y, p.z));
  auto x3 = f_sub(F, f_dbl(D));
  auto y3 = f_sub(f_mul(E, f_sub(D, x3)), f_dbl(f_dbl(f_dbl(C))));
  r.x = x3; r.y = y3; r.z = z3;
}
// mixed addition, q affine and finite; handles p = inf, p = q (doubling), p = -q (infinity)
template <typename J, typename A>
BN_JAC void jac_add_aff(J& r, const J& p, const A& q) {
  if (jac_is_inf(p)) { r.x = q.x; r.y = q.y; f_set_one(r.z); return; }
  auto z2 = f_sqr(p.z);
  auto u2 = f_mul(q.x, z2);
  auto s2 = f_mul(f_mul(q.y, z2), p.z);
  auto h = f_sub(u2, p.x);
  auto rr = f_sub(s2, p.y);
  if (f_is_zero(h)) {
    if (f_is_zero(rr)) { jac_dbl(r, p); return; }
    f_set_zero(r.x); f_set_zero(r.y); f_set_zero(r.z); return;
  }
  auto h2 = f_sqr(h), h3 = f_mul(h2, h), v = f_mul(p.x, h2);
  auto x3 = f_sub(f_sub(f_sqr(rr), h3), f_dbl(v));
  auto y3 = f_sub(f_mul(rr, f_sub(v, x3)), f_mul(p.y, h3));
  auto z3 = f_mul(p.z, h);
  r.x = x3; r.y = y3; r.z = z3;
}
template <typename J, typename A>
BN_HD void jac_to_aff(A& r, const J& p) {
  if (jac_is_inf(p)) { f_set_zero(r.x); f_set_zero(r.y); return; }
  auto zi = f_inv(p.z);
  auto zi2 = f_sqr(zi);
  r.x = f_mul(p.x, zi2);
  r.y = f_mul(p.y, f_mul(zi2, zi));
}
// the same with the inversion policy INV; no early exit (InvCta needs every thread): inv(0) = 0 maps infinity to (0, 0)
template <typename INV, typename J, typename A>
BN_HD void jac_to_aff_inv(A& r, const J& p) {
  auto zi = INV::inv(p.z);
  auto zi2 = f_sqr(zi);
  r.x = f_mul(p.x, zi2);
  r.y = f_mul(p.y, f_mul(zi2, zi));
}
template <typename A>
BN_HD bool aff_is_inf(const A& p) { return f_is_zero(p.x) && f_is_zero(p.y); }

// affine + affine with gnark Add semantics (infinity operands, doubling, P + (-P))
template <typename J, typename A>
BN_HD void aff_add(A& out, const A& a, const A& b) {
  if (aff_is_inf(a)) { out = b; return; }
  if (aff_is_inf(b)) { out = a; return; }
  J t; t.x = a.x; t.y = a.y; f_set_one(t.z);
  jac_add_aff(t, t, b);
  jac_to_aff(out, t);
}

// Jacobian + Jacobian (add-2007-bl) with the same edge handling as jac_add_aff
template <typename J>
BN_HD void jac_add(J& r, const J& p, const J& q) {
  if (jac_is_inf(p)) { r = q; return; }
  if (jac_is_inf(q)) { r = p; return; }
  auto z1z1 = f_sqr(p.z), z2z2 = f_sqr(q.z);
  auto u1 = f_mul(p.x, z2z2), u2 = f_mul(q.x, z1z1);
  auto s1 = f_mul(f_mul(p.y, q.z), z2z2), s2 = f_mul(f_mul(q.y, p.z), z1z1);
  auto h = f_sub(u2, u1), rr = f_sub(s2, s1);
  if (f_is_zero(h)) {
    if (f_is_zero(rr)) { jac_dbl(r, p); return; }
    f_set_zero(r.x); f_set_zero(r.y); f_set_zero(r.z); return;
  }
  auto hh = f_sqr(h), hhh = f_mul(hh, h), v = f_mul(u1, hh);
  auto x3 = f_sub(f_sub(f_sqr(rr), hhh), f_dbl(v));
  auto y3 = f_sub(f_mul(rr, f_sub(v, x3)), f_mul(s1, hhh));
  auto z3 = f_mul(f_mul(p.z, q.z), h);
  r.x = x3; r.y = y3; r.z = z3;
}

// ---- GLV: k = k1 + k2*lambda (mod r), |k1|,|k2| < 2^GLV_MAX_BITS, for ANY
// 256-bit k.  c_i = floor(k * G_iC / 2^256); k1 = k + c1*K1_M1 + c2*K1_M2, k2 = c1*K2_M1 + c2*K2_M2
// (mod 2^256, two's complement); the constants and their self-check live in gen_constants.py.
BN_HD void u256_mulhi(uint32_t* out, const uint32_t* a, const uint32_t* b) {
  uint32_t t[16];
  for (int i = 0; i < 16; i++) t[i] = 0;
  for (int i = 0; i < 8; i++) {
    uint64_t c = 0;
    for (int j = 0; j < 8; j++) { c += (uint64_t)a[i] * b[j] + t[i + j]; t[i + j] = (uint32_t)c; c >>= 32; }
    t[i + 8] = (uint32_t)c;
  }
  for (int i = 0; i < 8; i++) out[i] = t[i + 8];
}
BN_HD void u256_mullo_acc(uint32_t* acc, const uint32_t* a, const uint32_t* b) {  // acc += a*b mod 2^256
  for (int i = 0; i < 8; i++) {
    uint64_t c = 0;
    for (int j = 0; i + j < 8; j++) { c += (uint64_t)a[i] * b[j] + acc[i + j]; acc[i + j] = (uint32_t)c; c >>= 32; }
  }
}
BN_HD bool u256_abs(uint32_t* x) {  // two's complement -> magnitude; returns true when negative
  if (!(x[7] >> 31)) return false;
  uint64_t c = 1;
  for (int i = 0; i < 8; i++) { c += (uint32_t)~x[i]; x[i] = (uint32_t)c; c >>= 32; }
  return true;
}
BN_HD void glv_decompose(const uint32_t* k, uint32_t* k1, bool& neg1, uint32_t* k2, bool& neg2) {
  uint32_t g1[8], g2[8], m11[8], m12[8], m21[8], m22[8], c1[8], c2[8];
  for (int i = 0; i < 8; i++) { g1[i] = GLV_G1C[i]; g2[i] = GLV_G2C[i]; m11[i] = GLV_K1_M1[i]; m12[i] = GLV_K1_M2[i];
                                m21[i] = GLV_K2_M1[i]; m22[i] = GLV_K2_M2[i]; }
  u256_mulhi(c1, k, g1);
  u256_mulhi(c2, k, g2);
  for (int i = 0; i < 8; i++) { k1[i] = k[i]; k2[i] = 0; }
  u256_mullo_acc(k1, c1, m11); u256_mullo_acc(k1, c2, m12);
  u256_mullo_acc(k2, c1, m21); u256_mullo_acc(k2, c2, m22);
  neg1 = u256_abs(k1);
  neg2 = u256_abs(k2);
}
BN_HD Fp f_mul_beta(const Fp& x, const Fp& beta) { return f_mul(x, beta); }
BN_HD Fp2 f_mul_beta(const Fp2& x, const Fp& beta) { return fp2_mul_fp(x, beta); }

// [s]base by 2-dimensional GLV with the joint (Shamir) ladder over {P1, P2, P1+P2}; canonical affine out.
template <typename J, typename A, typename INV = InvThread>
BN_HD void scalar_mul_glv(A& out, const A& base, const uint32_t* s, const Fp& beta) {
  if (aff_is_inf(base)) { out = base; return; }  // (InvCta callers substitute a finite point: every thread must reach both inversions)
  uint32_t k1[8], k2[8];
  bool n1, n2;
  glv_decompose(s, k1, n1, k2, n2);
  A p1 = base, p2;
  p2.x = f_mul_beta(base.x, beta); p2.y = base.y;
  if (n1) p1.y = f_neg(p1.y);
  if (n2) p2.y = f_neg(p2.y);
  // table {P1, P2, P1+P2} in AFFINE form, indexed by the joint bit pair: every lane of a warp runs the SAME
  // doubling + one MIXED addition per bit (a data-dependent choice between addition paths would make each warp
  // execute all of them).  Normalising P1+P2 costs one inversion (~300 Fp-mul) and makes all ~96 additions of the
  // ladder mixed ones (G1: 11 instead of 16 Fp-mul each, G2: 30 instead of 44): -7 % / -18 % multiplications, and a
  // third less table on the stack.
  A tab[3];
  tab[0] = p1; tab[1] = p2;
  {
    J t; t.x = p1.x; t.y = p1.y; f_set_one(t.z);
    jac_add_aff(t, t, p2);
    jac_to_aff_inv<INV>(tab[2], t);  // (0, 0) if P1 + P2 is the point at infinity (only off the prime-order subgroup)
  }
  J acc; f_set_zero(acc.x); f_set_zero(acc.y); f_set_zero(acc.z);
  for (int i = GLV_MAX_BITS - 1; i >= 0; i--) {
    BN_CTA_SYNC();  // every thread runs all GLV_MAX_BITS iterations: a full, infinity-free CTA stays in lockstep
    jac_dbl(acc, acc);
    int b = (int)((k1[i >> 5] >> (i & 31)) & 1u) | ((int)((k2[i >> 5] >> (i & 31)) & 1u) << 1);
    if (b) {
      A e = tab[b - 1];
      if (!aff_is_inf(e)) jac_add_aff(acc, acc, e);
    }
  }
  jac_to_aff_inv<INV>(out, acc);
}

// ---- 4-dimensional GLS scalar multiplication on G2 (Galbraith-Scott) ------------------------------------------------
// psi = twist . Frobenius . untwist acts on the order-r subgroup G2 as multiplication by p mod r = 6 x0^2, a root of
// X^4 - X^2 + 1, so k = k0 + k1 psi + k2 psi^2 + k3 psi^3 with |k_j| < 2^GLS4_MAX_BITS (67 bits instead of the 130 of
// the 2-dimensional GLV split): HALF the doublings, and one mixed addition per bit from the 15-entry table of subset
// sums of (+-P, +-psi P, +-psi^2 P, +-psi^3 P).  ~4 400 Fp-mul instead of ~5 800.
// Valid for points of G2 proper -- every G2 point the reference produces (generator multiples, hash outputs with the
// cofactor cleared, sums of those); like gnark's own GLV ladder it is not scalar multiplication on other twist points.
BN_HD void g2_psi(G2Aff& r, const G2Aff& q) {
  r.x = f_mul(fp2_conj(q.x), GAMMA1[2]);
  r.y = f_mul(fp2_conj(q.y), GAMMA1[3]);
}
BN_HD void u128_mullo_acc(uint32_t* acc, const uint32_t* a, const uint32_t* b) {  // acc += a*b mod 2^128 (4 limbs)
  for (int i = 0; i < 4; i++) {
    uint64_t c = 0;
    for (int j = 0; i + j < 4; j++) { c += (uint64_t)a[i] * b[j] + acc[i + j]; acc[i + j] = (uint32_t)c; c >>= 32; }
  }
}
BN_HD void gls4_decompose(const uint32_t* k, uint32_t kv[4][4], bool neg[4]) {
  // c_i = floor(k * G_i / 2^256); k_j = [j == 0] k + sum_i c_i M_ij (mod 2^128, two's complement: |k_j| < 2^67)
  const uint32_t* G[4] = {GLS4_G0, GLS4_G1, GLS4_G2, GLS4_G3};
  const uint32_t* M[4][4] = {{GLS4_M00, GLS4_M01, GLS4_M02, GLS4_M03}, {GLS4_M10, GLS4_M11, GLS4_M12, GLS4_M13},
                             {GLS4_M20, GLS4_M21, GLS4_M22, GLS4_M23}, {GLS4_M30, GLS4_M31, GLS4_M32, GLS4_M33}};
  uint32_t c[4][8];
  for (int i = 0; i < 4; i++) {
    uint32_t g[8];
    for (int t = 0; t < 8; t++) g[t] = G[i][t];
    u256_mulhi(c[i], k, g);
  }
  for (int j = 0; j < 4; j++) {
    uint32_t acc[4];
    for (int t = 0; t < 4; t++) acc[t] = j == 0 ? k[t] : 0u;
    for (int i = 0; i < 4; i++) {
      uint32_t m[4];
      for (int t = 0; t < 4; t++) m[t] = M[i][j][t];
      u128_mullo_acc(acc, c[i], m);
    }
    neg[j] = (acc[3] >> 31) != 0;
    if (neg[j]) {
      uint64_t cy = 1;
      for (int t = 0; t < 4; t++) { cy += (uint32_t)~acc[t]; acc[t] = (uint32_t)cy; cy >>= 32; }
    }
    for (int t = 0; t < 4; t++) kv[j][t] = acc[t];
  }
}
// Per-thread scratch in GLOBAL memory (a contiguous slice: a per-lane table index on the local stack would touch 32
// sectors per word): 15 affine table entries, then 11 Jacobian Z coordinates and 11 prefix products of the batch inversion.
constexpr int kGlsTable = 15, kGlsComposite = 11;
constexpr int kGlsSliceFp2 = kGlsTable * 2 + 2 * kGlsComposite;  // in Fp2 units (64 B): 52 -> 3 328 B per thread
BN_HD G2Aff gls_ld(const Fp2* slice, int e) { G2Aff r; r.x = fp2_ld(slice[2 * e]); r.y = fp2_ld(slice[2 * e + 1]); return r; }
BN_HD void gls_st(Fp2* slice, int e, const G2Aff& v) { fp2_st(slice[2 * e], v.x); fp2_st(slice[2 * e + 1], v.y); }
template <typename INV = InvThread>
BN_HD void scalar_mul_gls4(G2Aff& out, const G2Aff& base, const uint32_t* s, Fp2* slice) {
  if (aff_is_inf(base)) { out = base; return; }  // (InvCta callers substitute a finite point)
  uint32_t kv[4][4];
  bool neg[4];
  gls4_decompose(s, kv, neg);
  Fp2* zs = slice + 2 * kGlsTable;
  Fp2* pre = zs + kGlsComposite;
  {
    G2Aff pj = base;
    for (int j = 0; j < 4; j++) {
      if (j) { G2Aff t; g2_psi(t, pj); pj = t; }
      G2Aff e = pj;
      if (neg[j]) e.y = f_neg(e.y);
      gls_st(slice, (1 << j) - 1, e);  // entry idx - 1 for idx = 1, 2, 4, 8
    }
  }
  // composite entries idx = rest + lowest bit, Jacobian for now: X, Y in the table slot, Z in zs[], prefix products in pre[]
  Fp2 run = fp2_one();
  int nc = 0;
  for (int idx = 3; idx < 16; idx++) {
    int low = idx & -idx, rest = idx ^ low;
    if (!rest) continue;
    G2Jac t;
    if (rest & (rest - 1)) {  // rest is itself composite (computed earlier: rest < idx)
      int rc = 0;
      for (int q = 3; q < rest; q++) if (q & (q - 1)) rc++;
      G2Aff xy = gls_ld(slice, rest - 1);
      t.x = xy.x; t.y = xy.y; t.z = fp2_ld(zs[rc]);
    } else {
      G2Aff xy = gls_ld(slice, rest - 1);
      t.x = xy.x; t.y = xy.y; t.z = fp2_one();
    }
    G2Aff pl = gls_ld(slice, low - 1);
    jac_add_aff(t, t, pl);  // handles doubling / cancellation (only off the subgroup)
    bool inf = jac_is_inf(t);
    G2Aff xy; xy.x = t.x; xy.y = t.y;
    Fp2 z = t.z;
    if (inf) { xy.x = fp2_zero(); xy.y = fp2_zero(); z = fp2_one(); }
    gls_st(slice, idx - 1, xy);
    fp2_st(zs[nc], z);
    run = f_mul(run, z);
    fp2_st(pre[nc], run);
    nc++;
  }
  {
    Fp2 inv = INV::inv(run);
    int ci = kGlsComposite - 1;
    for (int idx = 15; idx >= 3; idx--) {
      if (!(idx & (idx - 1))) continue;
      Fp2 zi = inv;
      if (ci > 0) zi = f_mul(inv, fp2_ld(pre[ci - 1]));
      inv = f_mul(inv, fp2_ld(zs[ci]));
      Fp2 zi2 = f_sqr(zi);
      G2Aff e = gls_ld(slice, idx - 1);
      e.x = f_mul(e.x, zi2);
      e.y = f_mul(e.y, f_mul(zi2, zi));
      gls_st(slice, idx - 1, e);  // (0, 0) stays (0, 0)
      ci--;
    }
  }
  G2Jac acc; f_set_zero(acc.x); f_set_zero(acc.y); f_set_zero(acc.z);
  for (int i = GLS4_MAX_BITS - 1; i >= 0; i--) {
    BN_CTA_SYNC();  // every thread runs all iterations: a full, infinity-free CTA stays in lockstep
    jac_dbl(acc, acc);
    int w = i >> 5, b = i & 31;
    int idx = (int)((kv[0][w] >> b) & 1u) | ((int)((kv[1][w] >> b) & 1u) << 1) | ((int)((kv[2][w] >> b) & 1u) << 2) | ((int)((kv[3][w] >> b) & 1u) << 3);
    if (idx) {
      G2Aff e = gls_ld(slice, idx - 1);
      if (!aff_is_inf(e)) jac_add_aff(acc, acc, e);
    }
  }
  jac_to_aff_inv<INV>(out, acc);
}

// Fixed base: table[w*255 + d-1] = [d * 2^(8w)] base (affine), w = 0..31, d = 1..255; 32 mixed additions.
constexpr int kFixedWindows = 32, kFixedEntries = 255;
template <typename J, typename A>
BN_HD void scalar_mul_fixed_jac(J& acc, const A* table, const uint32_t* s) {
  f_set_zero(acc.x); f_set_zero(acc.y); f_set_zero(acc.z);
  for (int w = 0; w < kFixedWindows; w++) {
    int d = (int)((s[w >> 2] >> ((w & 3) * 8)) & 0xFFu);
    if (d) {
      A e = table[w * kFixedEntries + d - 1];
      if (!aff_is_inf(e)) jac_add_aff(acc, acc, e);
    }
  }
}
template <typename J, typename A, typename INV = InvThread>
BN_HD void scalar_mul_fixed(A& out, const A* table, const uint32_t* s) {
  J acc;
  scalar_mul_fixed_jac<J, A>(acc, table, s);
  jac_to_aff_inv<INV>(out, acc);
}
// N Jacobian points -> affine with ONE inversion (Montgomery's trick over the finite z-coordinates): 3 (N - 1) extra
// products instead of N - 1 inversions of ~380 products each.  Same canonical affine results as jac_to_aff.
template <int N, typename J, typename A>
BN_HD void jac_to_aff_batch(A* out, const J* p, int count) {
  decltype(p[0].z) pre[N], run;
  f_set_one(run);
#pragma unroll
  for (int i = 0; i < N; i++) {
    pre[i] = run;
    if (i < count && !jac_is_inf(p[i])) run = f_mul(run, p[i].z);
  }
  auto inv = f_inv(run);
#pragma unroll
  for (int i = N - 1; i >= 0; i--) {
    if (i >= count) continue;
    if (jac_is_inf(p[i])) { f_set_zero(out[i].x); f_set_zero(out[i].y); continue; }
    auto zi = f_mul(inv, pre[i]);
    inv = f_mul(inv, p[i].z);
    auto zi2 = f_sqr(zi);
    out[i].x = f_mul(p[i].x, zi2);
    out[i].y = f_mul(p[i].y, f_mul(zi2, zi));
  }
}

// GT.Exp, generic Fp12 (no subgroup assumption), k = 256-bit LE, k == 0 -> 1.  FIXED 2-bit windows (gnark's
// E12.Exp shape): every lane of a warp multiplies at the same 128 positions, so the SIMT lanes never diverge on
// the exponent bits -- a bit-serial or sliding-window ladder would make every warp pay for the union of its
// lanes' multiplications.
// tab: 4 Fp12 of table space private to the thread (see gt_cyclo_exp about where it should live)
BN_HD void gt_exp(Fp12& out, const Fp12& x, const uint32_t* k, Fp12* tab) {
  // Control flow is identical for every thread (always two squarings and one product per window, the product by
  // tab[0] = 1 when the digit is 0), so a full CTA can run in lockstep; only the operand selection is per thread.
  // tab = 1, x, x^2, x^3
  fp12_set_one(tab[0]);
  tab[1] = x;
  fp12_sqr(tab[2], x);
  fp12_mul(tab[3], tab[2], x);
  Fp12 acc;
  fp12_set_one(acc);
  for (int w = 127; w >= 0; w--) {
    if (w != 127) { fp12_sqr(acc, acc); fp12_sqr(acc, acc); }
    int d = (int)((k[w >> 4] >> ((w & 15) * 2)) & 3u);
    Fp12 m = tab[d];
    fp12_mul(acc, acc, m);
  }
  out = acc;
}

// GT.Exp for bases in the cyclotomic subgroup (every pairing output and any product / quotient / power of
// pairing outputs -- all GT.Exp call sites of the reference): Granger-Scott squarings, FIXED
// 3-bit signed windows (digits -4..3, inverse = conjugate), table x..x^4: 258 cyclotomic squarings + <= 86
// products at lane-uniform positions.
BN_HD void gt_cyclo_exp_windowed(Fp12& out, const Fp12& x, const uint32_t* k) {
  Fp12 tab[4];  // x, x^2, x^3, x^4
  tab[0] = x;
  fp12_cyclo_sqr(tab[1], x);
  fp12_mul(tab[2], tab[1], x);
  fp12_cyclo_sqr(tab[3], tab[1]);
  // recode: k = sum d_i 8^i with d_i in [-4, 3]; 86 digits cover 258 bits (a carry may reach digit 85)
  signed char dg[87];
  int carry = 0;
  for (int i = 0; i < 86; i++) {
    int bit = 3 * i;
    int v = carry;
    if (bit < 256) {
      uint32_t lo = k[bit >> 5] >> (bit & 31);
      if ((bit & 31) > 29 && (bit >> 5) + 1 < 8) lo |= k[(bit >> 5) + 1] << (32 - (bit & 31));
      v += (int)(lo & 7u);  // at bit 255 only one bit exists: lo = k[7] >> 31
    }
    if (v > 3) { v -= 8; carry = 1; } else carry = 0;
    dg[i] = (signed char)v;
  }
  // uniform control flow (see gt_exp): three squarings and one product per digit, digit 0 multiplies by 1
  Fp12 acc;
  fp12_set_one(acc);
  for (int i = 85; i >= 0; i--) {
    if (i != 85) fp12_cyclo_sqr_n(acc, acc, 3);
    int d = dg[i];
    Fp12 m;
    if (d == 0) fp12_set_one(m);
    else if (d > 0) m = tab[d - 1];
    else fp12_conj(m, tab[-d - 1]);
    fp12_mul(acc, acc, m);
  }
  out = acc;
}

// GT.Exp for elements of GT proper (order r: pairing outputs and their products / quotients / powers) by the
// 2-dimensional GLV decomposition of the exponent that the G1 / G2 ladders use: p^2 = -lambda (mod r), so
// x^lambda = conj(frobenius^2(x)) costs 5 Fp2 x Fp products, and x^k = x^k1 * (x^lambda)^k2 with |k1|, |k2| < 2^129.
// Joint FIXED 2-bit windows over the 16-entry table x^i * (x^lambda)^j: 130 cyclotomic squarings + 65 products at
// lane-uniform positions + 11 products for the table, ~6.4 k Fp-mul instead of ~9.3 k for the 256-bit signed-window
// ladder above (kept as -DBN254_GT_EXP_WINDOWED; it also serves cyclotomic elements outside GT).
// tab: 16 Fp12 of caller-provided table space, private to the thread and CONTIGUOUS (the kernel hands out a slice of a
// global scratch: a per-thread digit indexes the table, and on the local-memory stack -- interleaved word by word
// across the lanes of a warp -- such a gather touches 32 sectors per word: ncu showed 22 GB of DRAM reads per wave and
// the multiply pipe at 47 %; one contiguous 384-byte entry per lane moves 1/8 of that).
// Round 2 (default): the FOUR-dimensional split the G2 ladder uses (gls4_decompose).  In GT the p-power Frobenius IS
// exponentiation by p, and p = lambda (mod r) is the eigenvalue of psi on G2, so the same lattice gives
// x^k = prod_j frob^j(x)^(k_j) with |k_j| < 2^GLS4_MAX_BITS: 66 cyclotomic squarings + 67 products from the 16-entry table
// of subset products of the four bases frob^j(x)^(+-1) (inverse = conjugate) + 11 products and three Frobenius maps for
// the table -- ~5.4 k Fp-mul instead of ~6.4 k for the two-dimensional split below (kept as -DBN254_GT_EXP_GLV2).
// Lane-uniform: every thread multiplies once per bit (by tab[0] = 1 when its four bits are zero).
BN_HD void gt_cyclo_exp_gls4(Fp12& out, const Fp12& x, const uint32_t* k, Fp12* tab) {
  uint32_t kv[4][4];
  bool neg[4];
  gls4_decompose(k, kv, neg);
  fp12_set_one(tab[0]);
  for (int j = 0; j < 4; j++) {
    Fp12 b;
    if (j == 0) b = x; else fp12_frob(b, x, j);
    if (neg[j]) fp12_conj(b, b);
    tab[1 << j] = b;
  }
  for (int idx = 3; idx < 16; idx++) {
    int low = idx & -idx, rest = idx ^ low;
    if (!rest) continue;
    fp12_mul(tab[idx], tab[rest], tab[low]);
  }
  Fp12 acc;
  fp12_set_one(acc);
  for (int i = GLS4_MAX_BITS - 1; i >= 0; i--) {
    if (i != GLS4_MAX_BITS - 1) fp12_cyclo_sqr(acc, acc);
    int w = i >> 5, b = i & 31;
    int idx = (int)((kv[0][w] >> b) & 1u) | ((int)((kv[1][w] >> b) & 1u) << 1) | ((int)((kv[2][w] >> b) & 1u) << 2) | ((int)((kv[3][w] >> b) & 1u) << 3);
    Fp12 m = tab[idx];
    fp12_mul(acc, acc, m);
  }
  out = acc;
}
BN_HD void gt_cyclo_exp(Fp12& out, const Fp12& x, const uint32_t* k, Fp12* tab) {
#ifdef BN254_GT_EXP_WINDOWED
  gt_cyclo_exp_windowed(out, x, k);
#elif !defined(BN254_GT_EXP_GLV2)
  gt_cyclo_exp_gls4(out, x, k, tab);
#else
  uint32_t k1[8], k2[8];
  bool n1, n2;
  glv_decompose(k, k1, n1, k2, n2);
  // tab[4 j + i] = a^i b^j, a = x^(+-1), b = (x^lambda)^(+-1)
  fp12_set_one(tab[0]);
  if (n1) fp12_conj(tab[1], x); else tab[1] = x;
  fp12_frob(tab[4], x, 2);                       // x^(-lambda)
  if (!n2) fp12_conj(tab[4], tab[4]);
  fp12_cyclo_sqr(tab[2], tab[1]);
  fp12_mul(tab[3], tab[2], tab[1]);
  fp12_cyclo_sqr(tab[8], tab[4]);
  fp12_mul(tab[12], tab[8], tab[4]);
  for (int j = 1; j < 4; j++)
    for (int i = 1; i < 4; i++) fp12_mul(tab[4 * j + i], tab[4 * j], tab[i]);
  constexpr int kWin = (GLV_MAX_BITS + 2) / 2;  // 2-bit windows covering bits 0 .. GLV_MAX_BITS
  Fp12 acc;
  fp12_set_one(acc);
  for (int w = kWin - 1; w >= 0; w--) {
    if (w != kWin - 1) fp12_cyclo_sqr_n(acc, acc, 2);
    int bit = 2 * w;
    int d = (int)((k1[bit >> 5] >> (bit & 31)) & 3u) | ((int)((k2[bit >> 5] >> (bit & 31)) & 3u) << 2);
    Fp12 m = tab[d];
    fp12_mul(acc, acc, m);
  }
  out = acc;
#endif
}

}  // namespace bn254
