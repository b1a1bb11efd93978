// BN254 base field Fp on 8 x 32-bit limbs, Montgomery form R = 2^256 (byte-identical to gnark's
// fp.Element = [4]uint64 little-endian limbs).  Hand-written for sm_100a:
// the multiply is a row-wise CIOS split into even/odd-aligned accumulators so that every
// 32x32->64 partial product is ONE carry-chained mad.lo.cc/madc.hi.cc pair, which ptxas fuses into
// IMAD.WIDE.U32(.X) on the fmaheavy pipe; additions ride the separate ALU pipe (IADD3.X).
//
// Replaces (reference side): gnark-crypto v0.19.0 ecc/bn254/fp/element.go (amd64 assembly), reached
// from every bn254.* call site listed in include/bn254_b200.h.
//
// Host emulation: when compiled by a plain C++ compiler (no __CUDACC__) the PTX wrappers fall back to a C model of
// the carry flag.  That build exists ONLY so tests/ can exercise the device algorithms on a box
// without a GPU (tests/emu); the shipped library has no CPU path.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define BN_HD __device__ __forceinline__
#define BN_D __device__ __forceinline__
#define BN_NOINLINE static __device__ __noinline__
#else
#define BN_HD static inline
#define BN_D static inline
#define BN_NOINLINE static __attribute__((noinline))
struct uint4 { uint32_t x, y, z, w; };  // host-emulation stand-in for the CUDA vector type
#endif

#if defined(__CUDACC__)
#define BN_CONST static __device__ __constant__ const
#else
#define BN_CONST static const
#endif

namespace bn254 {

struct alignas(16) Fp { uint32_t l[8]; };
struct Fp2 { Fp a0, a1; };

// ---------------------------------------------------------------------------------------------
// PTX carry-chain primitives (one instruction each).  asm volatile keeps program order, so the CC
// flag set by one wrapper is the one read by the next.
#if defined(__CUDACC__)
BN_D uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t d; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
BN_D uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t d; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
BN_D uint32_t addc(uint32_t a, uint32_t b) { uint32_t d; asm volatile("addc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
BN_D uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t d; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
BN_D uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t d; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
BN_D uint32_t subc(uint32_t a, uint32_t b) { uint32_t d; asm volatile("subc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
BN_D uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
BN_D uint32_t mul_hi(uint32_t a, uint32_t b) { return __umulhi(a, b); }
BN_D uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t d; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
BN_D uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t d; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
BN_D uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t d; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
BN_D uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t d; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
#else
static thread_local uint32_t bn_cc = 0;  // emulated PTX condition-code register (carry / NOT borrow handled below)
static inline uint32_t add_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b; bn_cc = (uint32_t)(t >> 32); return (uint32_t)t; }
static inline uint32_t addc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b + bn_cc; bn_cc = (uint32_t)(t >> 32); return (uint32_t)t; }
static inline uint32_t addc(uint32_t a, uint32_t b) { return a + b + bn_cc; }
// PTX: sub.cc writes the borrow into CC.CF; subc computes a - b - CF.
static inline uint32_t sub_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b; bn_cc = (uint32_t)((t >> 32) & 1); return (uint32_t)t; }
static inline uint32_t subc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b - bn_cc; bn_cc = (uint32_t)((t >> 32) & 1); return (uint32_t)t; }
static inline uint32_t subc(uint32_t a, uint32_t b) { return a - b - bn_cc; }
static inline uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
static inline uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (uint64_t)(uint32_t)(a * b) + c; bn_cc = (uint32_t)(t >> 32); return (uint32_t)t; }
static inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (uint64_t)(uint32_t)(a * b) + c + bn_cc; bn_cc = (uint32_t)(t >> 32); return (uint32_t)t; }
static inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (((uint64_t)a * b) >> 32) + c + bn_cc; bn_cc = (uint32_t)(t >> 32); return (uint32_t)t; }
static inline uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return (uint32_t)((((uint64_t)a * b) >> 32) + c + bn_cc); }
#endif

}  // namespace bn254

#include "bn254_constants.cuh"

namespace bn254 {

BN_HD uint32_t p_limb(int i) {
  switch (i) { case 0: return P0; case 1: return P1; case 2: return P2; case 3: return P3;
               case 4: return P4; case 5: return P5; case 6: return P6; default: return P7; }
}

BN_HD Fp fp_zero() { Fp z; for (int i = 0; i < 8; i++) z.l[i] = 0; return z; }
BN_HD Fp fp_one() { Fp z = BN254_FP_ONE; return z; }
BN_HD bool fp_is_zero(const Fp& a) { uint32_t o = 0; for (int i = 0; i < 8; i++) o |= a.l[i]; return o == 0; }
BN_HD bool fp_eq(const Fp& a, const Fp& b) { uint32_t o = 0; for (int i = 0; i < 8; i++) o |= a.l[i] ^ b.l[i]; return o == 0; }

// t (< 2p) -> t mod p : subtract p, keep the difference unless it borrowed.
BN_HD void fp_reduce_once(Fp& t) {
  uint32_t d[8];
  d[0] = sub_cc(t.l[0], P0); d[1] = subc_cc(t.l[1], P1); d[2] = subc_cc(t.l[2], P2); d[3] = subc_cc(t.l[3], P3);
  d[4] = subc_cc(t.l[4], P4); d[5] = subc_cc(t.l[5], P5); d[6] = subc_cc(t.l[6], P6); d[7] = subc_cc(t.l[7], P7);
  uint32_t borrow = subc(0u, 0u);  // 0 or 0xffffffff
#pragma unroll
  for (int i = 0; i < 8; i++) t.l[i] = borrow ? t.l[i] : d[i];
}

BN_HD Fp fp_add(const Fp& a, const Fp& b) {
  Fp t;
  t.l[0] = add_cc(a.l[0], b.l[0]); t.l[1] = addc_cc(a.l[1], b.l[1]); t.l[2] = addc_cc(a.l[2], b.l[2]); t.l[3] = addc_cc(a.l[3], b.l[3]);
  t.l[4] = addc_cc(a.l[4], b.l[4]); t.l[5] = addc_cc(a.l[5], b.l[5]); t.l[6] = addc_cc(a.l[6], b.l[6]); t.l[7] = addc(a.l[7], b.l[7]);
  fp_reduce_once(t);  // a,b < p < 2^254: the sum fits 8 limbs
  return t;
}
BN_HD Fp fp_dbl(const Fp& a) { return fp_add(a, a); }

BN_HD Fp fp_sub(const Fp& a, const Fp& b) {
  Fp t;
  t.l[0] = sub_cc(a.l[0], b.l[0]); t.l[1] = subc_cc(a.l[1], b.l[1]); t.l[2] = subc_cc(a.l[2], b.l[2]); t.l[3] = subc_cc(a.l[3], b.l[3]);
  t.l[4] = subc_cc(a.l[4], b.l[4]); t.l[5] = subc_cc(a.l[5], b.l[5]); t.l[6] = subc_cc(a.l[6], b.l[6]); t.l[7] = subc_cc(a.l[7], b.l[7]);
  uint32_t borrow = subc(0u, 0u);  // all-ones iff a < b
#if defined(__CUDACC__)
  // add p back under a predicate: 8 predicated adds instead of 8 masks + 8 adds
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %8, 0;\n\t"
      "@q add.cc.u32 %0, %0, %9;\n\t@q addc.cc.u32 %1, %1, %10;\n\t@q addc.cc.u32 %2, %2, %11;\n\t@q addc.cc.u32 %3, %3, %12;\n\t"
      "@q addc.cc.u32 %4, %4, %13;\n\t@q addc.cc.u32 %5, %5, %14;\n\t@q addc.cc.u32 %6, %6, %15;\n\t@q addc.u32 %7, %7, %16;\n\t}"
      : "+r"(t.l[0]), "+r"(t.l[1]), "+r"(t.l[2]), "+r"(t.l[3]), "+r"(t.l[4]), "+r"(t.l[5]), "+r"(t.l[6]), "+r"(t.l[7])
      : "r"(borrow), "r"(P0), "r"(P1), "r"(P2), "r"(P3), "r"(P4), "r"(P5), "r"(P6), "r"(P7));
#else
  t.l[0] = add_cc(t.l[0], P0 & borrow); t.l[1] = addc_cc(t.l[1], P1 & borrow); t.l[2] = addc_cc(t.l[2], P2 & borrow); t.l[3] = addc_cc(t.l[3], P3 & borrow);
  t.l[4] = addc_cc(t.l[4], P4 & borrow); t.l[5] = addc_cc(t.l[5], P5 & borrow); t.l[6] = addc_cc(t.l[6], P6 & borrow); t.l[7] = addc(t.l[7], P7 & borrow);
#endif
  return t;
}
BN_HD Fp fp_neg(const Fp& a) { return fp_sub(fp_zero(), a); }  // -0 = 0 (borrow never set)

// a/2 mod p
BN_HD Fp fp_half(const Fp& a) {
  uint32_t odd = 0u - (a.l[0] & 1u);
  uint32_t t[9];
  t[0] = add_cc(a.l[0], P0 & odd); t[1] = addc_cc(a.l[1], P1 & odd); t[2] = addc_cc(a.l[2], P2 & odd); t[3] = addc_cc(a.l[3], P3 & odd);
  t[4] = addc_cc(a.l[4], P4 & odd); t[5] = addc_cc(a.l[5], P5 & odd); t[6] = addc_cc(a.l[6], P6 & odd); t[7] = addc_cc(a.l[7], P7 & odd);
  t[8] = addc(0u, 0u);
  Fp z;
#pragma unroll
  for (int i = 0; i < 8; i++) z.l[i] = (t[i] >> 1) | (t[i + 1] << 31);
  return z;
}

// ---- Montgomery product -----------------------------------------------------------------------
// Accumulator split: ev[j] holds word position j (pairs (0,1),(2,3),..), od[j] holds word position
// j+1 (pairs (1,2),(3,4),..).  Products a[even]*bi land on ev pairs, a[odd]*bi on od pairs, so each
// row is two independent 8-instruction carry chains.  After the reduction row ev[0] == 0; dropping
// it swaps the roles of the two arrays (the old ev[1] is folded into the new ev[0]).
BN_HD void mont_row_first(uint32_t* ev, uint32_t* od, const uint32_t* a, uint32_t bi) {
#pragma unroll
  for (int j = 0; j < 8; j += 2) {
    ev[j] = mul_lo(a[j], bi); ev[j + 1] = mul_hi(a[j], bi);
    od[j] = mul_lo(a[j + 1], bi); od[j + 1] = mul_hi(a[j + 1], bi);
  }
}
BN_HD void mont_row_next(uint32_t* ev, uint32_t* od, const uint32_t* a, uint32_t bi) {
  // 'od' is the previous row's even array: od[0] == 0 is dropped, od[1] folds into ev[0], od[2..7] shift down
  ev[0] = add_cc(ev[0], od[1]);
  od[0] = madc_lo_cc(a[1], bi, od[2]); od[1] = madc_hi_cc(a[1], bi, od[3]);
  od[2] = madc_lo_cc(a[3], bi, od[4]); od[3] = madc_hi_cc(a[3], bi, od[5]);
  od[4] = madc_lo_cc(a[5], bi, od[6]); od[5] = madc_hi_cc(a[5], bi, od[7]);
  od[6] = madc_lo_cc(a[7], bi, 0u);    od[7] = madc_hi(a[7], bi, 0u);
  ev[0] = mad_lo_cc(a[0], bi, ev[0]);  ev[1] = madc_hi_cc(a[0], bi, ev[1]);
  ev[2] = madc_lo_cc(a[2], bi, ev[2]); ev[3] = madc_hi_cc(a[2], bi, ev[3]);
  ev[4] = madc_lo_cc(a[4], bi, ev[4]); ev[5] = madc_hi_cc(a[4], bi, ev[5]);
  ev[6] = madc_lo_cc(a[6], bi, ev[6]); ev[7] = madc_hi_cc(a[6], bi, ev[7]);
  od[7] = addc(od[7], 0u);
}
#if defined(__CUDACC__)
// Same instruction sequence as mont_row_next, emitted as ONE asm block: ptxas then fuses every
// mad.lo.cc/madc.hi.cc pair of the row into IMAD.WIDE.U32.X (with separate asm statements it splits
// the a*b rows into IMAD + IMAD.HI + 2 IADD3.X).  Which form is faster on sm_100a is measured by
// profiles/microbench/imad_peak.cu; BN254_MUL_VARIANT selects it.
BN_D void mont_row_next_fused(uint32_t* ev, uint32_t* od, const uint32_t* a, uint32_t bi) {
  asm volatile(
      "add.cc.u32 %0, %0, %9;\n\t"
      "madc.lo.cc.u32 %8, %16, %20, %10;\n\t madc.hi.cc.u32 %9, %16, %20, %11;\n\t"
      "madc.lo.cc.u32 %10, %17, %20, %12;\n\t madc.hi.cc.u32 %11, %17, %20, %13;\n\t"
      "madc.lo.cc.u32 %12, %18, %20, %14;\n\t madc.hi.cc.u32 %13, %18, %20, %15;\n\t"
      "madc.lo.cc.u32 %14, %19, %20, 0;\n\t madc.hi.u32 %15, %19, %20, 0;\n\t"
      "mad.lo.cc.u32 %0, %21, %20, %0;\n\t madc.hi.cc.u32 %1, %21, %20, %1;\n\t"
      "madc.lo.cc.u32 %2, %22, %20, %2;\n\t madc.hi.cc.u32 %3, %22, %20, %3;\n\t"
      "madc.lo.cc.u32 %4, %23, %20, %4;\n\t madc.hi.cc.u32 %5, %23, %20, %5;\n\t"
      "madc.lo.cc.u32 %6, %24, %20, %6;\n\t madc.hi.cc.u32 %7, %24, %20, %7;\n\t"
      "addc.u32 %15, %15, 0;"
      : "+r"(ev[0]), "+r"(ev[1]), "+r"(ev[2]), "+r"(ev[3]), "+r"(ev[4]), "+r"(ev[5]), "+r"(ev[6]), "+r"(ev[7]),
        "+r"(od[0]), "+r"(od[1]), "+r"(od[2]), "+r"(od[3]), "+r"(od[4]), "+r"(od[5]), "+r"(od[6]), "+r"(od[7])
      : "r"(a[1]), "r"(a[3]), "r"(a[5]), "r"(a[7]), "r"(bi), "r"(a[0]), "r"(a[2]), "r"(a[4]), "r"(a[6]));
}
#else
static inline void mont_row_next_fused(uint32_t* ev, uint32_t* od, const uint32_t* a, uint32_t bi) { mont_row_next(ev, od, a, bi); }
#endif
#ifndef BN254_MUL_VARIANT
#define BN254_MUL_VARIANT 1  // 0: per-instruction asm (ptxas splits a*b rows), 1: fused IMAD.WIDE.X rows
#endif
template <int V>
BN_HD void mont_row_next_v(uint32_t* ev, uint32_t* od, const uint32_t* a, uint32_t bi) {
  if (V == 0) mont_row_next(ev, od, a, bi); else mont_row_next_fused(ev, od, a, bi);
}
BN_HD void mont_row_reduce(uint32_t* ev, uint32_t* od) {
  uint32_t m = ev[0] * P_INV32;
  od[0] = mad_lo_cc(P1, m, od[0]);  od[1] = madc_hi_cc(P1, m, od[1]);
  od[2] = madc_lo_cc(P3, m, od[2]); od[3] = madc_hi_cc(P3, m, od[3]);
  od[4] = madc_lo_cc(P5, m, od[4]); od[5] = madc_hi_cc(P5, m, od[5]);
  od[6] = madc_lo_cc(P7, m, od[6]); od[7] = madc_hi(P7, m, od[7]);
  ev[0] = mad_lo_cc(P0, m, ev[0]);  ev[1] = madc_hi_cc(P0, m, ev[1]);
  ev[2] = madc_lo_cc(P2, m, ev[2]); ev[3] = madc_hi_cc(P2, m, ev[3]);
  ev[4] = madc_lo_cc(P4, m, ev[4]); ev[5] = madc_hi_cc(P4, m, ev[5]);
  ev[6] = madc_lo_cc(P6, m, ev[6]); ev[7] = madc_hi_cc(P6, m, ev[7]);
  od[7] = addc(od[7], 0u);
}

// z = a*b/R mod p, canonical.  Requires a, b < 2p (verified by tests/emu): every intermediate row then
// fits the 9-word even/odd accumulators and the value before the final conditional subtraction is < 2p.
template <int V>
BN_HD Fp fp_mul_v(const Fp& a, const Fp& b) {
  uint32_t e[8], o[8];
  mont_row_first(e, o, a.l, b.l[0]); mont_row_reduce(e, o);
  mont_row_next_v<V>(o, e, a.l, b.l[1]);  mont_row_reduce(o, e);
  mont_row_next_v<V>(e, o, a.l, b.l[2]);  mont_row_reduce(e, o);
  mont_row_next_v<V>(o, e, a.l, b.l[3]);  mont_row_reduce(o, e);
  mont_row_next_v<V>(e, o, a.l, b.l[4]);  mont_row_reduce(e, o);
  mont_row_next_v<V>(o, e, a.l, b.l[5]);  mont_row_reduce(o, e);
  mont_row_next_v<V>(e, o, a.l, b.l[6]);  mont_row_reduce(e, o);
  mont_row_next_v<V>(o, e, a.l, b.l[7]);  mont_row_reduce(o, e);
  // now o is even-aligned with o[0] == 0; result word k = e[k] + o[k+1]
  Fp z;
  z.l[0] = add_cc(e[0], o[1]); z.l[1] = addc_cc(e[1], o[2]); z.l[2] = addc_cc(e[2], o[3]); z.l[3] = addc_cc(e[3], o[4]);
  z.l[4] = addc_cc(e[4], o[5]); z.l[5] = addc_cc(e[5], o[6]); z.l[6] = addc_cc(e[6], o[7]); z.l[7] = addc(e[7], 0u);
  fp_reduce_once(z);
  return z;
}
BN_HD Fp fp_mul(const Fp& a, const Fp& b) { return fp_mul_v<BN254_MUL_VARIANT>(a, b); }

BN_HD Fp fp_sqr(const Fp& a) { return fp_mul(a, a); }

// ---- lazy reduction building blocks (Aranha et al. style): a wide 8x8 -> 16-limb product and a
// separate Montgomery reduction, so an Fp2 product costs 3 wide products + 2 reductions (320 IMAD.WIDE)
// instead of 3 full Montgomery products (384).
// x[0..7] += (a0,a1,a2,a3) * b laid on the four aligned word pairs, carry out added into cw.
BN_HD void mac4(uint32_t* x, uint32_t& cw, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b) {
#if defined(__CUDACC__)
  asm volatile(
      "mad.lo.cc.u32 %0, %9, %13, %0;\n\t madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
      "madc.lo.cc.u32 %2, %10, %13, %2;\n\t madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
      "madc.lo.cc.u32 %4, %11, %13, %4;\n\t madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
      "madc.lo.cc.u32 %6, %12, %13, %6;\n\t madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
      "addc.u32 %8, %8, 0;"
      : "+r"(x[0]), "+r"(x[1]), "+r"(x[2]), "+r"(x[3]), "+r"(x[4]), "+r"(x[5]), "+r"(x[6]), "+r"(x[7]), "+r"(cw)
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
#else
  x[0] = mad_lo_cc(a0, b, x[0]);  x[1] = madc_hi_cc(a0, b, x[1]);
  x[2] = madc_lo_cc(a1, b, x[2]); x[3] = madc_hi_cc(a1, b, x[3]);
  x[4] = madc_lo_cc(a2, b, x[4]); x[5] = madc_hi_cc(a2, b, x[5]);
  x[6] = madc_lo_cc(a3, b, x[6]); x[7] = madc_hi_cc(a3, b, x[7]);
  cw = addc(cw, 0u);
#endif
}
// t[0..15] = a * b (plain integer product of the limb vectors; a, b < 2^256)
BN_HD void fp_mul_wide(uint32_t* t, const Fp& a, const Fp& b) {
  uint32_t E[18], O[18];  // E[k]: word position k (aligned pairs (0,1),(2,3)..); O[k]: word position k+1
#pragma unroll
  for (int i = 0; i < 18; i++) { E[i] = 0; O[i] = 0; }
#pragma unroll
  for (int i = 0; i < 8; i += 2) {
    // even multiplier word b[i]: a-even products sit on E pairs at i, a-odd products on O pairs at i
    mac4(&E[i], E[i + 8], a.l[0], a.l[2], a.l[4], a.l[6], b.l[i]);
    mac4(&O[i], O[i + 8], a.l[1], a.l[3], a.l[5], a.l[7], b.l[i]);
    // odd multiplier word b[i+1]: a-odd products at position i+1+j (even) -> E pairs at i+2; a-even -> O pairs at i
    mac4(&E[i + 2], E[i + 10], a.l[1], a.l[3], a.l[5], a.l[7], b.l[i + 1]);
    mac4(&O[i], O[i + 8], a.l[0], a.l[2], a.l[4], a.l[6], b.l[i + 1]);
  }
  t[0] = E[0];
  t[1] = add_cc(E[1], O[0]);
#pragma unroll
  for (int k = 2; k < 15; k++) t[k] = addc_cc(E[k], O[k - 1]);
  t[15] = addc(E[15], O[14]);
}
// one row of the windowed Montgomery reduction: frame shifts right by one word (ev <-> od swap roles),
// the next input word t_in enters at frame position 7, and m*p is added.  See fp_redc.
BN_HD void redc_row(uint32_t* ev, uint32_t* od, uint32_t t_in) {
#if defined(__CUDACC__)
  asm volatile(
      "{\n\t.reg .u32 m;\n\t"
      "add.cc.u32 %0, %0, %9;\n\t"
      "mul.lo.u32 m, %0, %25;\n\t"
      "madc.lo.cc.u32 %8, %18, m, %10;\n\t madc.hi.cc.u32 %9, %18, m, %11;\n\t"
      "madc.lo.cc.u32 %10, %20, m, %12;\n\t madc.hi.cc.u32 %11, %20, m, %13;\n\t"
      "madc.lo.cc.u32 %12, %22, m, %14;\n\t madc.hi.cc.u32 %13, %22, m, %15;\n\t"
      "madc.lo.cc.u32 %14, %24, m, %16;\n\t madc.hi.u32 %15, %24, m, 0;\n\t"
      "mad.lo.cc.u32 %0, %17, m, %0;\n\t madc.hi.cc.u32 %1, %17, m, %1;\n\t"
      "madc.lo.cc.u32 %2, %19, m, %2;\n\t madc.hi.cc.u32 %3, %19, m, %3;\n\t"
      "madc.lo.cc.u32 %4, %21, m, %4;\n\t madc.hi.cc.u32 %5, %21, m, %5;\n\t"
      "madc.lo.cc.u32 %6, %23, m, %6;\n\t madc.hi.cc.u32 %7, %23, m, %7;\n\t"
      "addc.u32 %15, %15, 0;\n\t}"
      : "+r"(ev[0]), "+r"(ev[1]), "+r"(ev[2]), "+r"(ev[3]), "+r"(ev[4]), "+r"(ev[5]), "+r"(ev[6]), "+r"(ev[7]),
        "+r"(od[0]), "+r"(od[1]), "+r"(od[2]), "+r"(od[3]), "+r"(od[4]), "+r"(od[5]), "+r"(od[6]), "+r"(od[7])
      : "r"(t_in), "r"(P0), "r"(P1), "r"(P2), "r"(P3), "r"(P4), "r"(P5), "r"(P6), "r"(P7), "r"(P_INV32));
#else
  ev[0] = add_cc(ev[0], od[1]);
  uint32_t m = ev[0] * P_INV32;
  od[0] = madc_lo_cc(P1, m, od[2]); od[1] = madc_hi_cc(P1, m, od[3]);
  od[2] = madc_lo_cc(P3, m, od[4]); od[3] = madc_hi_cc(P3, m, od[5]);
  od[4] = madc_lo_cc(P5, m, od[6]); od[5] = madc_hi_cc(P5, m, od[7]);
  od[6] = madc_lo_cc(P7, m, t_in);  od[7] = madc_hi(P7, m, 0u);
  ev[0] = mad_lo_cc(P0, m, ev[0]);  ev[1] = madc_hi_cc(P0, m, ev[1]);
  ev[2] = madc_lo_cc(P2, m, ev[2]); ev[3] = madc_hi_cc(P2, m, ev[3]);
  ev[4] = madc_lo_cc(P4, m, ev[4]); ev[5] = madc_hi_cc(P4, m, ev[5]);
  ev[6] = madc_lo_cc(P6, m, ev[6]); ev[7] = madc_hi_cc(P6, m, ev[7]);
  od[7] = addc(od[7], 0u);
#endif
}
// z = t / R mod p, canonical, for a 16-limb t < p * 2^256.
BN_HD Fp fp_redc(const uint32_t* t) {
  uint32_t e[8], o[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { e[i] = t[i]; o[i] = 0; }
  mont_row_reduce(e, o);          // frame 0: e positions 0..7, o positions 1..8
  redc_row(o, e, t[8]);           // frame 1
  redc_row(e, o, t[9]);
  redc_row(o, e, t[10]);
  redc_row(e, o, t[11]);
  redc_row(o, e, t[12]);
  redc_row(e, o, t[13]);
  redc_row(o, e, t[14]);          // frame 7: o is even-aligned with o[0] == 0
  Fp z;
  z.l[0] = add_cc(e[0], o[1]); z.l[1] = addc_cc(e[1], o[2]); z.l[2] = addc_cc(e[2], o[3]); z.l[3] = addc_cc(e[3], o[4]);
  z.l[4] = addc_cc(e[4], o[5]); z.l[5] = addc_cc(e[5], o[6]); z.l[6] = addc_cc(e[6], o[7]); z.l[7] = addc(e[7], t[15]);
  fp_reduce_once(z);
  return z;
}
// 16-limb helpers
BN_HD void wide_sub(uint32_t* z, const uint32_t* a, const uint32_t* b, uint32_t& borrow_mask) {
  z[0] = sub_cc(a[0], b[0]);
#pragma unroll
  for (int i = 1; i < 16; i++) z[i] = subc_cc(a[i], b[i]);
  borrow_mask = subc(0u, 0u);
}
BN_HD uint32_t psq_limb(int i) {
  switch (i) { case 0: return PSQ0; case 1: return PSQ1; case 2: return PSQ2; case 3: return PSQ3; case 4: return PSQ4;
               case 5: return PSQ5; case 6: return PSQ6; case 7: return PSQ7; case 8: return PSQ8; case 9: return PSQ9;
               case 10: return PSQ10; case 11: return PSQ11; case 12: return PSQ12; case 13: return PSQ13;
               case 14: return PSQ14; default: return PSQ15; }
}
// z += p^2 & mask
BN_HD void wide_add_psq_masked(uint32_t* z, uint32_t mask) {
  z[0] = add_cc(z[0], PSQ0 & mask);
#pragma unroll
  for (int i = 1; i < 15; i++) z[i] = addc_cc(z[i], psq_limb(i) & mask);
  z[15] = addc(z[15], PSQ15 & mask);
}
// a + b without reduction (a, b < p: fits 8 limbs since 2p < 2^255)
BN_HD Fp fp_add_noreduce(const Fp& a, const Fp& b) {
  Fp t;
  t.l[0] = add_cc(a.l[0], b.l[0]); t.l[1] = addc_cc(a.l[1], b.l[1]); t.l[2] = addc_cc(a.l[2], b.l[2]); t.l[3] = addc_cc(a.l[3], b.l[3]);
  t.l[4] = addc_cc(a.l[4], b.l[4]); t.l[5] = addc_cc(a.l[5], b.l[5]); t.l[6] = addc_cc(a.l[6], b.l[6]); t.l[7] = addc(a.l[7], b.l[7]);
  return t;
}

// a^(p-2); inv(0) = 0 like gnark's Inverse.
BN_HD Fp fp_inv(const Fp& a) {
  Fp acc = fp_one(), b = a;
  for (int i = 0; i < 254; i++) {
    if ((FP_PM2[i >> 5] >> (i & 31)) & 1u) acc = fp_mul(acc, b);
    b = fp_sqr(b);
  }
  return acc;
}

}  // namespace bn254
