// Extension tower over Fp for BN254, in gnark's E2/E6/E12 memory order:
//   Fp2 = Fp[u]/(u^2+1), Fp6 = Fp2[v]/(v^3-(9+u)), Fp12 = Fp6[w]/(w^2-v).
// Replaces (reference side): gnark-crypto v0.19.0 ecc/bn254/internal/fptower/{e2,e6,e12,e12_pairing,
// frobenius}.go, reached through GT.Mul/Div/Exp (access/tree/access_tree_node.go:114,156-157;
// cpabe/bsw07/bsw07_cpabe.go:189-190) and every bn254.Pair call.
//
// Layering: Fp2 arithmetic is register-level (one thread, 16+16 limbs live); Fp6/Fp12 values live in
// memory (per-thread local or shared) and are combined by out-of-line routines built from Fp2 steps,
// which keeps the instruction footprint small enough for the SM instruction caches.
#pragma once
#include "fp.cuh"

namespace bn254 {

struct Fp6 { Fp2 b0, b1, b2; };
struct Fp12 { Fp6 c0, c1; };

// 128-bit operand access for the out-of-line leaves: their reference parameters are generic pointers (local
// stack or the shared-memory scratch); word-wise generic loads cost 4x the LSU instructions and 4-way bank
// conflicts on the 592-byte scratch stride.
#if defined(__CUDACC__)
BN_D Fp2 fp2_ld(const Fp2& r) {
  Fp2 v;
  const uint4* p = reinterpret_cast<const uint4*>(&r);
  uint4* d = reinterpret_cast<uint4*>(&v);
  d[0] = p[0]; d[1] = p[1]; d[2] = p[2]; d[3] = p[3];
  return v;
}
BN_D void fp2_st(Fp2& r, const Fp2& v) {
  uint4* p = reinterpret_cast<uint4*>(&r);
  const uint4* d = reinterpret_cast<const uint4*>(&v);
  p[0] = d[0]; p[1] = d[1]; p[2] = d[2]; p[3] = d[3];
}
BN_D Fp fp_ld(const Fp& r) { Fp v; const uint4* p = reinterpret_cast<const uint4*>(&r); uint4* d = reinterpret_cast<uint4*>(&v); d[0] = p[0]; d[1] = p[1]; return v; }
#else
BN_D Fp2 fp2_ld(const Fp2& r) { return r; }
BN_D void fp2_st(Fp2& r, const Fp2& v) { r = v; }
BN_D Fp fp_ld(const Fp& r) { return r; }
#endif
// Add-type Fp2 leaves: inline by default; -DBN254_OOL_ADDS makes them out-of-line calls, trading
// call overhead for a much smaller instruction footprint (ncu: k_pair v1 stalls on no_instruction).
#ifdef BN254_OOL_ADDS
#define BN_LEAF BN_NOINLINE
#else
#define BN_LEAF BN_HD
#endif
// -DBN254_OOL_FPMUL: the Montgomery product itself becomes ONE out-of-line body (operands and result
// passed by value in registers), so fp2_mul / fp2_sqr / fp2_mul_fp shrink to a few calls.
#ifdef BN254_OOL_FPMUL
BN_NOINLINE Fp fp_mul_call(Fp a, Fp b) { return fp_mul(a, b); }
#define FP_MUL(a, b) fp_mul_call(a, b)
#else
#define FP_MUL(a, b) fp_mul(a, b)
#endif
// ------------------------------------------------------------------------------------------ Fp2
BN_HD Fp2 fp2_zero() { Fp2 z; z.a0 = fp_zero(); z.a1 = fp_zero(); return z; }
BN_HD Fp2 fp2_one() { Fp2 z; z.a0 = fp_one(); z.a1 = fp_zero(); return z; }
BN_HD bool fp2_is_zero(const Fp2& a) { return fp_is_zero(a.a0) && fp_is_zero(a.a1); }
BN_HD bool fp2_eq(const Fp2& a, const Fp2& b) { return fp_eq(a.a0, b.a0) && fp_eq(a.a1, b.a1); }
// _i = always-inline bodies (used by the tower VM, where each op exists once); the unsuffixed names are
// the BN_LEAF versions used by the one-thread-per-pairing routines.
BN_HD Fp2 fp2_add_i(const Fp2& a, const Fp2& b) { Fp2 z; z.a0 = fp_add(a.a0, b.a0); z.a1 = fp_add(a.a1, b.a1); return z; }
BN_HD Fp2 fp2_sub_i(const Fp2& a, const Fp2& b) { Fp2 z; z.a0 = fp_sub(a.a0, b.a0); z.a1 = fp_sub(a.a1, b.a1); return z; }
BN_HD Fp2 fp2_dbl_i(const Fp2& a) { Fp2 z; z.a0 = fp_dbl(a.a0); z.a1 = fp_dbl(a.a1); return z; }
BN_HD Fp2 fp2_neg_i(const Fp2& a) { Fp2 z; z.a0 = fp_neg(a.a0); z.a1 = fp_neg(a.a1); return z; }
BN_HD Fp2 fp2_conj_i(const Fp2& a) { Fp2 z; z.a0 = a.a0; z.a1 = fp_neg(a.a1); return z; }
BN_HD Fp2 fp2_half_i(const Fp2& a) { Fp2 z; z.a0 = fp_half(a.a0); z.a1 = fp_half(a.a1); return z; }
BN_HD Fp2 fp2_mul_fp_i(const Fp2& a, const Fp& k) { Fp2 z; z.a0 = FP_MUL(a.a0, k); z.a1 = FP_MUL(a.a1, k); return z; }
// (9x + y) mod p for canonical x and y <= p.  9x + y < 10p is formed in 9 limbs (8x by funnel shifts), the quotient
// by p is estimated from the top 13 bits (never too large, at most 1 too small: checked exhaustively at the
// boundaries and on 2M random values by csrc/gen_constants.py's derivation, tests/test_emu_device_code.py), q*p is
// subtracted and one conditional subtraction finishes.  ~70 instructions instead of ~125 for three reduced
// doublings and two reduced additions; the 8 multiplies by q run while the multiply pipe is otherwise idle
// (xi sits in the add-type phase of the leaves).
// v (9 limbs, 0 <= v < 64p) -> v mod p, canonical: quotient estimate from the top bits, one multiple of p subtracted,
// one conditional subtraction.
BN_HD Fp fp_reduce_small(const uint32_t* v) {
  uint32_t q = ((((v[8] << 11) | (v[7] >> 21)) * 43336u) >> 24);  // floor(v / 2^245) * floor(2^269 / p) >> 24: q or q - 1
  uint32_t qp[9];
  uint64_t c = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) { c += (uint64_t)q * p_limb(i); qp[i] = (uint32_t)c; c >>= 32; }
  qp[8] = (uint32_t)c;
  Fp w;
  w.l[0] = sub_cc(v[0], qp[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) w.l[i] = subc_cc(v[i], qp[i]);
  (void)subc(v[8], qp[8]);  // the ninth limb of v - q p is zero: the difference is below 2p
  fp_reduce_once(w);
  return w;
}
BN_HD uint32_t bn_shl3(uint32_t lo, uint32_t hi) { return (hi << 3) | (lo >> 29); }
BN_HD Fp fp_mul9_add(const Fp& x, const uint32_t* y) {
  uint32_t v[9];
  v[0] = add_cc(x.l[0] << 3, x.l[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) v[i] = addc_cc(bn_shl3(x.l[i - 1], x.l[i]), x.l[i]);
  v[8] = addc(x.l[7] >> 29, 0u);
  v[0] = add_cc(v[0], y[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) v[i] = addc_cc(v[i], y[i]);
  v[8] = addc(v[8], 0u);
  return fp_reduce_small(v);
}
// (9+u)(a0 + a1 u) = (9a0 - a1) + (a0 + 9a1) u
BN_HD Fp2 fp2_mul_xi_i(const Fp2& a) {
#ifndef BN254_XI_DOUBLINGS
  uint32_t na1[8];  // p - a1 (in (0, p]; no borrow)
  na1[0] = sub_cc(P0, a.a1.l[0]); na1[1] = subc_cc(P1, a.a1.l[1]); na1[2] = subc_cc(P2, a.a1.l[2]); na1[3] = subc_cc(P3, a.a1.l[3]);
  na1[4] = subc_cc(P4, a.a1.l[4]); na1[5] = subc_cc(P5, a.a1.l[5]); na1[6] = subc_cc(P6, a.a1.l[6]); na1[7] = subc(P7, a.a1.l[7]);
  Fp2 z;
  z.a0 = fp_mul9_add(a.a0, na1);
  z.a1 = fp_mul9_add(a.a1, a.a0.l);
  return z;
#else
  Fp e0 = fp_dbl(fp_dbl(fp_dbl(a.a0))), e1 = fp_dbl(fp_dbl(fp_dbl(a.a1)));
  Fp2 z;
  z.a0 = fp_sub(fp_add(e0, a.a0), a.a1);
  z.a1 = fp_add(fp_add(e1, a.a1), a.a0);
  return z;
#endif
}
BN_LEAF Fp2 fp2_add(const Fp2& a, const Fp2& b) { return fp2_add_i(fp2_ld(a), fp2_ld(b)); }
BN_LEAF Fp2 fp2_sub(const Fp2& a, const Fp2& b) { return fp2_sub_i(fp2_ld(a), fp2_ld(b)); }
BN_LEAF Fp2 fp2_dbl(const Fp2& a) { return fp2_dbl_i(fp2_ld(a)); }
BN_LEAF Fp2 fp2_neg(const Fp2& a) { return fp2_neg_i(fp2_ld(a)); }
BN_LEAF Fp2 fp2_conj(const Fp2& a) { return fp2_conj_i(fp2_ld(a)); }
BN_LEAF Fp2 fp2_half(const Fp2& a) { return fp2_half_i(fp2_ld(a)); }
BN_LEAF Fp2 fp2_mul_fp(const Fp2& a, const Fp& k) { return fp2_mul_fp_i(fp2_ld(a), fp_ld(k)); }
BN_LEAF Fp2 fp2_mul_xi(const Fp2& a) { return fp2_mul_xi_i(fp2_ld(a)); }
// Karatsuba: 3 Fp products
BN_HD Fp2 fp2_mul_inl(const Fp2& a, const Fp2& b) {
  Fp t0 = FP_MUL(a.a0, b.a0);
  Fp t1 = FP_MUL(a.a1, b.a1);
  Fp m = FP_MUL(fp_add_noreduce(a.a0, a.a1), fp_add_noreduce(b.a0, b.a1));  // operands < 2p are fine for the product
  Fp2 z;
  z.a0 = fp_sub(t0, t1);
  z.a1 = fp_sub(fp_sub(m, t0), t1);
  return z;
}
// Lazy-reduction Karatsuba (Aranha et al.): 3 wide products, 2 Montgomery reductions -- 320 IMAD.WIDE instead of
// the 384 of three full Montgomery products.
//   c1 = (a0+a1)(b0+b1) - a0b0 - a1b1 >= 0 ;  c0 = a0b0 - a1b1 (+ p^2 if negative) ; both < 2p^2 < p*2^256
BN_HD Fp2 fp2_mul_lazy(const Fp2& a, const Fp2& b) {
  uint32_t t0[16], t1[16], t2[16];
  fp_mul_wide(t0, a.a0, b.a0);
  fp_mul_wide(t1, a.a1, b.a1);
  fp_mul_wide(t2, fp_add_noreduce(a.a0, a.a1), fp_add_noreduce(b.a0, b.a1));
  uint32_t mask;
  wide_sub(t2, t2, t0, mask);
  wide_sub(t2, t2, t1, mask);
  wide_sub(t0, t0, t1, mask);
  wide_add_psq_masked(t0, mask);
  Fp2 z;
  z.a0 = fp_redc(t0);
  z.a1 = fp_redc(t2);
  return z;
}
// The same arithmetic in an order that keeps fewer values alive (experiment for a 128-register build: two wide products,
// their difference and sum, THEN the third product on the operand sums): c1 = (a0+a1)(b0+b1) - (a0b0 + a1b1).
BN_HD Fp2 fp2_mul_lazy_seq(const Fp2& a, const Fp2& b) {
  uint32_t t0[16], t1[16], s[16];
  fp_mul_wide(t0, a.a0, b.a0);
  fp_mul_wide(t1, a.a1, b.a1);
  s[0] = add_cc(t0[0], t1[0]);
#pragma unroll
  for (int i = 1; i < 15; i++) s[i] = addc_cc(t0[i], t1[i]);
  s[15] = addc(t0[15], t1[15]);  // a0b0 + a1b1 < 2 p^2 < 2^509
  uint32_t mask;
  wide_sub(t0, t0, t1, mask);
  wide_add_psq_masked(t0, mask);
  fp_mul_wide(t1, fp_add_noreduce(a.a0, a.a1), fp_add_noreduce(b.a0, b.a1));
  wide_sub(t1, t1, s, mask);
  Fp2 z;
  z.a0 = fp_redc(t0);
  z.a1 = fp_redc(t1);
  return z;
}
// complex squaring: 2 Fp products
BN_HD Fp2 fp2_sqr_inl(const Fp2& a) {
  Fp m = FP_MUL(a.a0, a.a1);
  Fp2 z;
  z.a0 = FP_MUL(fp_add_noreduce(a.a0, a.a1), fp_sub(a.a0, a.a1));
  z.a1 = fp_dbl(m);
  return z;
}
BN_HD Fp2 fp2_mul_best(const Fp2& a, const Fp2& b) { return fp2_mul_inl(a, b); }  // inline form (tower VM)
// The Fp2 product of the thread kernels: ONE out-of-line body, operands and result in registers, lazy reduction
// inside, so none of the 16-limb intermediates crosses a call boundary (as separate out-of-line wide-product /
// reduction calls the same arithmetic measured 10-20 % SLOWER than three Montgomery products; fused it is
// 5.5 % faster: 2.83M vs 2.69M pairings/s).  -DBN254_KARATSUBA_MULX restores the three-product body.
#if defined(BN254_MULX_SEQ)
BN_NOINLINE Fp2 fp2_mulx(Fp2 a, Fp2 b) { return fp2_mul_lazy_seq(a, b); }
#elif !defined(BN254_KARATSUBA_MULX)
BN_NOINLINE Fp2 fp2_mulx(Fp2 a, Fp2 b) { return fp2_mul_lazy(a, b); }
#else
BN_NOINLINE Fp2 fp2_mulx(Fp2 a, Fp2 b) { return fp2_mul_inl(a, b); }
#endif
#define FP2_MUL(a, b) fp2_mulx(a, b)
#define FP2_SQR(a) fp2_sqr_inl(a)
// out-of-line bodies shared by every tower routine
BN_NOINLINE void fp2_mul(Fp2& z, const Fp2& a, const Fp2& b) { fp2_st(z, FP2_MUL(fp2_ld(a), fp2_ld(b))); }
BN_NOINLINE void fp2_sqr(Fp2& z, const Fp2& a) { fp2_st(z, FP2_SQR(fp2_ld(a))); }
BN_NOINLINE void fp_inv_ool(Fp& z, const Fp& a) { z = fp_inv(a); }
BN_HD void fp2_inv(Fp2& z, const Fp2& a) {
  Fp n = fp_add(fp_sqr(a.a0), fp_sqr(a.a1));
  Fp ni; fp_inv_ool(ni, n);
  Fp2 r; r.a0 = fp_mul(a.a0, ni); r.a1 = fp_neg(fp_mul(a.a1, ni));
  z = r;
}

// -DBN254_SMEM_SCRATCH: the Fp2 temporaries of the innermost composite routines (fp6_mul, fp6_mul_01,
// fp12_cyclo_sqr) live in a per-thread slice of dynamic shared memory instead of the local-memory stack
// (ncu: the 7 KB/thread stack drives 3 TB/s of DRAM traffic and a third of all stall samples).  9 slots of
// 64 B, thread stride 592 B = 16 B x 37 (odd) so 128-bit accesses of a quarter-warp hit distinct bank quads.
constexpr int kScratchSlots = 9;
#ifndef BN254_SCRATCH_STRIDE
#define BN254_SCRATCH_STRIDE 592
#endif
constexpr int kScratchStride = BN254_SCRATCH_STRIDE;  // (a smaller stride is a TIMING-ONLY experiment: slots then overlap between threads)
#if defined(BN254_SMEM_SCRATCH) && defined(__CUDACC__)
extern __shared__ uint4 bn_dyn_smem[];
BN_D Fp2* bn_scratch() { return reinterpret_cast<Fp2*>(reinterpret_cast<char*>(bn_dyn_smem) + threadIdx.x * kScratchStride); }
#define BN_SCRATCH_DECL Fp2* sc_ = bn_scratch();
// out-of-line routines that receive the scratch pointer re-derive it so the compiler knows it is shared memory
#define BN_SC_REBIND(sc) sc = bn_scratch();
#else
#define BN_SCRATCH_DECL Fp2 sc_[kScratchSlots];
#define BN_SC_REBIND(sc)
#endif
// -DBN254_CTA_LOCKSTEP: the warps of a CTA run the same instruction stream (one pairing per thread, no
// data-dependent control flow), so keeping them loosely in step -- a CTA barrier at the entry of every composite
// routine -- lets the four warps share each instruction-cache line fetched into the SM's L1.5 (32 KB against a
// ~60 KB hot loop; ncu: no_instruction was the top stall of the staged build).  The barrier is taken only when the
// kernel has established that EVERY thread of the CTA follows the same path (bn_lockstep, CTA-uniform).
#if defined(BN254_CTA_LOCKSTEP) && defined(__CUDACC__)
__shared__ int bn_lockstep;
#define BN_CTA_SYNC() do { if (bn_lockstep) __syncthreads(); } while (0)
BN_D void cta_lockstep_set(bool uniform) { __syncthreads(); if (threadIdx.x == 0) bn_lockstep = uniform ? 1 : 0; __syncthreads(); }
BN_D bool cta_lockstep_on() { return bn_lockstep != 0; }
#else
#define BN_CTA_SYNC() do { } while (0)
BN_D void cta_lockstep_set(bool) {}
BN_D bool cta_lockstep_on() { return false; }
#endif
// ------------------------------------------------------------------------------------------ Fp6
BN_HD void fp6_add(Fp6& z, const Fp6& x, const Fp6& y) { z.b0 = fp2_add(x.b0, y.b0); z.b1 = fp2_add(x.b1, y.b1); z.b2 = fp2_add(x.b2, y.b2); }
BN_HD void fp6_sub(Fp6& z, const Fp6& x, const Fp6& y) { z.b0 = fp2_sub(x.b0, y.b0); z.b1 = fp2_sub(x.b1, y.b1); z.b2 = fp2_sub(x.b2, y.b2); }
BN_HD void fp6_neg(Fp6& z, const Fp6& x) { z.b0 = fp2_neg(x.b0); z.b1 = fp2_neg(x.b1); z.b2 = fp2_neg(x.b2); }
BN_HD void fp6_mul_v(Fp6& z, const Fp6& x) { Fp2 t = fp2_mul_xi(x.b2); z.b2 = x.b1; z.b1 = x.b0; z.b0 = t; }
BN_NOINLINE void fp6_mul(Fp6& z, const Fp6& x, const Fp6& y);  // tower_staged.cuh
BN_NOINLINE void fp6_inv(Fp6& z, const Fp6& x) {
  Fp2 t0, t1, t2, s, n;
  fp2_sqr(t0, x.b0); fp2_mul(s, x.b1, x.b2); t0 = fp2_sub(t0, fp2_mul_xi(s));
  fp2_sqr(t1, x.b2); fp2_mul(s, x.b0, x.b1); t1 = fp2_sub(fp2_mul_xi(t1), s);
  fp2_sqr(t2, x.b1); fp2_mul(s, x.b0, x.b2); t2 = fp2_sub(t2, s);
  fp2_mul(n, x.b2, t1); fp2_mul(s, x.b1, t2); n = fp2_mul_xi(fp2_add(n, s));
  fp2_mul(s, x.b0, t0); n = fp2_add(n, s);
  fp2_inv(n, n);
  Fp2 r0, r1, r2;
  fp2_mul(r0, t0, n); fp2_mul(r1, t1, n); fp2_mul(r2, t2, n);
  z.b0 = r0; z.b1 = r1; z.b2 = r2;
}

// ------------------------------------------------------------------------------------------ Fp12
BN_HD void fp12_set_one(Fp12& z) {
  z.c0.b0 = fp2_one(); z.c0.b1 = fp2_zero(); z.c0.b2 = fp2_zero();
  z.c1.b0 = fp2_zero(); z.c1.b1 = fp2_zero(); z.c1.b2 = fp2_zero();
}
BN_HD bool fp12_is_one(const Fp12& z) {
  return fp2_eq(z.c0.b0, fp2_one()) && fp2_is_zero(z.c0.b1) && fp2_is_zero(z.c0.b2) &&
         fp2_is_zero(z.c1.b0) && fp2_is_zero(z.c1.b1) && fp2_is_zero(z.c1.b2);
}
BN_NOINLINE void fp12_mul(Fp12& z, const Fp12& x, const Fp12& y);
BN_NOINLINE void fp12_sqr(Fp12& z, const Fp12& x);
BN_HD void fp12_conj(Fp12& z, const Fp12& x) { z.c0 = x.c0; fp6_neg(z.c1, x.c1); }
BN_NOINLINE void fp12_inv(Fp12& z, const Fp12& x) {
  Fp6 n, t;
  fp6_mul(n, x.c0, x.c0); fp6_mul(t, x.c1, x.c1); fp6_mul_v(t, t); fp6_sub(n, n, t);
  fp6_inv(n, n);
  fp6_mul(t, x.c1, n);
  fp6_mul(z.c0, x.c0, n);
  fp6_neg(z.c1, t);
}
// p^k-power Frobenius, k in {1,2,3}.  In the w-basis (g0=c0.b0 g1=c1.b0 g2=c0.b1 g3=c1.b1 g4=c0.b2
// g5=c1.b2) the map is g_i -> conj^k(g_i) * xi^(i (p^k-1)/6).
BN_NOINLINE void fp12_frob(Fp12& z, const Fp12& x, int k) {
  Fp2 g[6] = {x.c0.b0, x.c1.b0, x.c0.b1, x.c1.b1, x.c0.b2, x.c1.b2};
  if (k & 1) { for (int i = 0; i < 6; i++) g[i] = fp2_conj(g[i]); }
  for (int i = 1; i < 6; i++) {
    if (k == 2) g[i] = fp2_mul_fp(g[i], GAMMA2[i]);
    else { Fp2 c = (k == 1) ? GAMMA1[i] : GAMMA3[i]; fp2_mul(g[i], g[i], c); }
  }
  z.c0.b0 = g[0]; z.c1.b0 = g[1]; z.c0.b1 = g[2]; z.c1.b1 = g[3]; z.c0.b2 = g[4]; z.c1.b2 = g[5];
}
#include "tower_staged.cuh"
// x^e for x in the cyclotomic subgroup, e given as width-3 signed digits (LSB first); inverse = conjugate
BN_NOINLINE void fp12_cyclo_exp_naf3(Fp12& z, const Fp12& x, const signed char* digits, int len) {
  // (copying the table entry and conjugating the copy measured 6 % FASTER on B200 than folding the conjugation
  // into a second product routine: the extra routine costs more instruction-cache than the 384-byte copy)
  Fp12 x3, acc, m;
  fp12_cyclo_sqr(x3, x); fp12_mul(x3, x3, x);
  bool started = false;
  int pending = 0;  // squarings owed to acc: runs between non-zero digits are done in one staged pass
  for (int i = len - 1; i >= 0; i--) {
    if (started) pending++;
    int d = digits[i];
    if (d) {
      if (pending) { fp12_cyclo_sqr_n(acc, acc, pending); pending = 0; }
      int ad = d < 0 ? -d : d;
      if (ad == 1) m = x; else m = x3;
      if (d < 0) fp12_conj(m, m);
      if (started) fp12_mul(acc, acc, m); else { acc = m; started = true; }
    }
  }
  if (pending) fp12_cyclo_sqr_n(acc, acc, pending);
  z = acc;
}
BN_HD void fp12_expt(Fp12& z, const Fp12& x) { fp12_cyclo_exp_naf3(z, x, X0_NAF3, X0_NAF3_LEN); }

}  // namespace bn254
