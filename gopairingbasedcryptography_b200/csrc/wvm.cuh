// Warp-VM interpreter: ONE WARP per pairing, one Fp-level micro-op per lane per round (wvmgen.py).
// Every Fp value is a 32-byte slot in the warp's shared-memory slice.  Three op classes, one class per round:
//   MUL  d = (s0 +- s1) * (s2 +- s3)       operands canonical, the +- pre-additions unreduced (< 2p), Montgomery product
//   LIN  d = sum_i c_i * s_i  (|c_i| <= 31, sum |c_i| <= 120, <= 15 terms)   carry-free column sums, one small reduction
//   INV  d = s0^-1                         lane-local binary extended Euclid (one per final exponentiation)
// A slot read in round r is never written in round r (the allocator recycles a slot only after the round of its last
// use), so lanes need no ordering inside a round; one __syncwarp() separates rounds.
//
// Replaces (reference side): the same gnark calls as pairing.cuh -- bn254.Pair / MillerLoop / FinalExponentiation with
// the 1-element slices of the reference's 60 call sites and the small batches of BASELINE configs 0 and 4.
#pragma once
#include "tower.cuh"

namespace bn254 { namespace wvm {

enum : unsigned { OP_NOP = 0, OP_MUL = 1, OP_LIN = 2, OP_INV = 3 };
constexpr int kLanes = 32;

#if defined(__CUDACC__)
BN_D Fp ld_slot(const Fp* slots, unsigned s) { return fp_ld(slots[s]); }
BN_D void st_slot(Fp* slots, unsigned s, const Fp& v) {
  uint4* p = reinterpret_cast<uint4*>(&slots[s]);
  const uint4* d = reinterpret_cast<const uint4*>(&v);
  p[0] = d[0]; p[1] = d[1];
}
BN_D void round_sync() { __syncwarp(); }
BN_D unsigned warp_max(unsigned x) { return __reduce_max_sync(0xffffffffu, x); }
#else
BN_D Fp ld_slot(const Fp* slots, unsigned s) { return slots[s]; }
BN_D void st_slot(Fp* slots, unsigned s, const Fp& v) { slots[s] = v; }
BN_D void round_sync() {}
#endif

// s (canonical) -> s or p - s; p - s lies in [1, p], which the unreduced pre-addition of a MUL operand accepts
BN_HD Fp cond_neg(const Fp& s, bool neg) {
  Fp d;
  d.l[0] = sub_cc(P0, s.l[0]); d.l[1] = subc_cc(P1, s.l[1]); d.l[2] = subc_cc(P2, s.l[2]); d.l[3] = subc_cc(P3, s.l[3]);
  d.l[4] = subc_cc(P4, s.l[4]); d.l[5] = subc_cc(P5, s.l[5]); d.l[6] = subc_cc(P6, s.l[6]); d.l[7] = subc(P7, s.l[7]);
#pragma unroll
  for (int i = 0; i < 8; i++) d.l[i] = neg ? d.l[i] : s.l[i];
  return d;
}

// LIN accumulator: eight independent 64-bit column sums (no carry chain between limbs: a lone warp has nothing else
// to hide a dependent chain behind).  A negative term c * v enters as |c| * ~v (per-limb complement) and the count
// K = sum |c| over the negative terms is kept: sum |c| ~v = K (2^256 - 1) - sum |c| v, corrected once at the end.
// The accumulator starts at 128 p, so the total stays positive: 0 < 128 p + sum c_i s_i < 256 p  (sum |c_i| <= 120).
struct LinAcc { uint64_t col[8]; uint32_t k; };
BN_HD void lin_init(LinAcc& a) {
#pragma unroll
  for (int i = 0; i < 8; i++) a.col[i] = (uint64_t)p_limb(i) << 7;  // 128 p, limb by limb (carries resolved at the end)
  a.k = 0;
}
BN_HD void lin_acc(LinAcc& a, const Fp& v, int c) {
  const uint32_t mask = c < 0 ? 0xFFFFFFFFu : 0u;
  const uint32_t m = (uint32_t)(c < 0 ? -c : c);
#pragma unroll
  for (int i = 0; i < 8; i++) a.col[i] += (uint64_t)(v.l[i] ^ mask) * m;
  a.k += mask & m;
}
// One pass: quotient estimate from the two top columns (the lower columns cannot move it by more than one unit of
// 2^224, i.e. 2^-21 of the estimate's own unit), then columns - q p with signed 64-bit carries, then at most two
// conditional subtractions (q_est in {q - 2, q - 1, q}).
#ifndef WVM_FINISH_ONEPASS
#define WVM_FINISH_ONEPASS 0
#endif
#ifndef WVM_LIN_BATCH
#define WVM_LIN_BATCH 1  // LIN terms in batches of four loads: Pair(1) 1.44 -> 1.32 ms (profiles/r2/wvm_interpreter_variants.jsonl)
#endif
#ifndef WVM_AHEAD
#define WVM_AHEAD 1
#endif
#if WVM_FINISH_ONEPASS
BN_HD Fp lin_finish(const LinAcc& a) {
  uint64_t top = a.col[7] + (a.col[6] >> 32) - ((uint64_t)a.k << 32);     // floor(v / 2^224) up to the lower carries
  uint32_t hi = (uint32_t)(top >> 21);                                   // floor(v / 2^245) < 2^17
  uint32_t q = (uint32_t)(((uint64_t)hi * 43336u) >> 24);                // 43336 = floor(2^269 / p)
  Fp w;
  int64_t c = (int64_t)a.k;                                              // + K at limb 0 (see lin_acc)
#pragma unroll
  for (int i = 0; i < 8; i++) {
    c += (int64_t)a.col[i] - (int64_t)((uint64_t)q * p_limb(i));
    w.l[i] = (uint32_t)c;
    c >>= 32;                                                            // arithmetic: the running value may dip below zero
  }
  // c - K is the ninth limb of v - q p, 0 by construction (0 <= v - q p < 3p < 2^256)
  fp_reduce_once(w);
  fp_reduce_once(w);
  return w;
}
#else
// two passes: carry propagation of the columns (plus K, minus K 2^256), then the quotient estimate of fp_reduce_small
// (tower.cuh) with the product widened to 64 bits -- for v < 256 p, q_est is q or q - 1 -- and one conditional subtraction
BN_HD Fp lin_finish(const LinAcc& a) {
  uint32_t v[9];
  uint64_t c = a.k;
#pragma unroll
  for (int i = 0; i < 8; i++) { c += a.col[i]; v[i] = (uint32_t)c; c >>= 32; }
  v[8] = (uint32_t)c - a.k;
  uint32_t hi = (v[8] << 11) | (v[7] >> 21);
  uint32_t q = (uint32_t)(((uint64_t)hi * 43336u) >> 24);
  uint32_t qp[9];
  uint64_t d = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) { d += (uint64_t)q * p_limb(i); qp[i] = (uint32_t)d; d >>= 32; }
  qp[8] = (uint32_t)d;
  Fp w;
  w.l[0] = sub_cc(v[0], qp[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) w.l[i] = subc_cc(v[i], qp[i]);
  (void)subc(v[8], qp[8]);
  fp_reduce_once(w);
  return w;
}
#endif
BN_HD Fp lin_reduce9(const uint32_t* v) {  // test hook: reduce a plain 9-limb value below 256 p through the same path
  LinAcc a;
  for (int i = 0; i < 8; i++) a.col[i] = v[i];
  a.col[7] += (uint64_t)v[8] << 32;
  a.k = 0;
  return lin_finish(a);
}

// Inversion for the warp-VM: binary extended Euclid on one lane (variable time is free here -- the other lanes idle
// during the INV round either way) instead of the 380 dependent Montgomery products of the Fermat chain.
//   input a R (Montgomery form) -> (a R)^-1 mod p by Stein's algorithm -> times R^3 / R = a^-1 R.   inv(0) = 0.
BN_HD bool u256_is_one(const uint32_t* x) { uint32_t o = x[0] ^ 1u; for (int i = 1; i < 8; i++) o |= x[i]; return o == 0; }
BN_HD bool u256_geq(const uint32_t* a, const uint32_t* b) {
  for (int i = 7; i >= 0; i--) if (a[i] != b[i]) return a[i] > b[i];
  return true;
}
BN_HD void u256_sub(uint32_t* a, const uint32_t* b) {
  uint64_t bw = 0;
  for (int i = 0; i < 8; i++) { uint64_t d = (uint64_t)a[i] - b[i] - bw; a[i] = (uint32_t)d; bw = (d >> 32) & 1; }
}
BN_HD void u256_half_mod(uint32_t* x) {  // x / 2 mod p for x < p
  uint64_t c = 0;
  uint32_t t[9];
  uint32_t odd = 0u - (x[0] & 1u);
  for (int i = 0; i < 8; i++) { c += (uint64_t)x[i] + (p_limb(i) & odd); t[i] = (uint32_t)c; c >>= 32; }
  t[8] = (uint32_t)c;
  for (int i = 0; i < 8; i++) x[i] = (t[i] >> 1) | (t[i + 1] << 31);
}
BN_HD void u256_sub_mod(uint32_t* a, const uint32_t* b) {  // a - b mod p, a, b < p
  uint64_t bw = 0;
  for (int i = 0; i < 8; i++) { uint64_t d = (uint64_t)a[i] - b[i] - bw; a[i] = (uint32_t)d; bw = (d >> 32) & 1; }
  if (bw) { uint64_t c = 0; for (int i = 0; i < 8; i++) { c += (uint64_t)a[i] + p_limb(i); a[i] = (uint32_t)c; c >>= 32; } }
}
BN_NOINLINE Fp fp_inv_wvm(Fp a) {
  if (fp_is_zero(a)) return a;
  uint32_t u[8], v[8], x1[8], x2[8];
  for (int i = 0; i < 8; i++) { u[i] = a.l[i]; v[i] = p_limb(i); x1[i] = i == 0; x2[i] = 0; }
  while (!u256_is_one(u) && !u256_is_one(v)) {
    while (!(u[0] & 1u)) { for (int i = 0; i < 7; i++) u[i] = (u[i] >> 1) | (u[i + 1] << 31); u[7] >>= 1; u256_half_mod(x1); }
    while (!(v[0] & 1u)) { for (int i = 0; i < 7; i++) v[i] = (v[i] >> 1) | (v[i + 1] << 31); v[7] >>= 1; u256_half_mod(x2); }
    if (u256_geq(u, v)) { u256_sub(u, v); u256_sub_mod(x1, x2); }
    else { u256_sub(v, u); u256_sub_mod(x2, x1); }
  }
  Fp r;
  const uint32_t* x = u256_is_one(u) ? x1 : x2;
  for (int i = 0; i < 8; i++) r.l[i] = x[i];
  Fp r3 = BN254_FP_R3;
  return fp_mul(r, r3);  // (a R)^-1 * R^3 / R = a^-1 R
}

// (A Montgomery product on carry-free 64-bit column accumulators -- every partial product split into halves, no carry
// chain -- was measured for the lone-warp case and is SLOWER: 2 390 cycles per MUL round against 1 250 for fp_mul's two
// carry chains per row; its ~600 instructions are issue-bound on a single warp.  profiles/r2/wvm_round_costs.jsonl)
// One op.  rec = 16 u16 fields in two uint4 (little-endian pairs).  Returns true when `out` must be stored to slot dst.
// nmax: warp-uniform bound on the LIN term count of this round (so the term loop's exit is a uniform branch).
BN_HD unsigned field(const uint4& w0, const uint4& w1, int i) {
  const uint32_t ws[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
  return (ws[i >> 1] >> ((i & 1) * 16)) & 0xFFFFu;
}
BN_HD bool exec_op(const Fp* slots, const uint4& w0, const uint4& w1, unsigned nmax, unsigned& dst, Fp& out) {
  const unsigned h = field(w0, w1, 0);
  const unsigned op = h & 3u, x = h >> 12;
  dst = (h >> 2) & 1023u;
  if (op == OP_MUL) {
    Fp a = ld_slot(slots, field(w0, w1, 1)), a1 = ld_slot(slots, field(w0, w1, 2));
    Fp b = ld_slot(slots, field(w0, w1, 3)), b1 = ld_slot(slots, field(w0, w1, 4));
    a = fp_add_noreduce(a, cond_neg(a1, (x & 1u) != 0));
    b = fp_add_noreduce(b, cond_neg(b1, (x & 2u) != 0));
    out = fp_mul(a, b);
    return true;
  }
  if (op == OP_LIN) {
    // terms in batches of four: the four slot loads of a batch are in flight together, the batch loop's exit is
    // warp-uniform; unused term fields are zero = coefficient 0 on slot 0
    LinAcc acc;
    lin_init(acc);
    (void)x;
#if !WVM_LIN_BATCH
#pragma unroll
    for (int i = 0; i < 15; i++) {  // term by term; the exit is warp-uniform, unused fields are coefficient 0 on slot 0
      if ((unsigned)i >= nmax) break;
      unsigned f = field(w0, w1, 1 + i);
      int c = (int)(f >> 10);
      c = c >= 32 ? c - 64 : c;
      lin_acc(acc, ld_slot(slots, f & 1023u), c);
    }
#else
#pragma unroll
    for (int g = 0; g < 16; g += 4) {
      if ((unsigned)g >= nmax) break;
      Fp v[4];
      int c[4];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        unsigned f = g + i < 15 ? field(w0, w1, 1 + g + i) : 0u;
        c[i] = (int)(f >> 10);
        c[i] = c[i] >= 32 ? c[i] - 64 : c[i];
        v[i] = ld_slot(slots, f & 1023u);
      }
#pragma unroll
      for (int i = 0; i < 4; i++) lin_acc(acc, v[i], c[i]);
    }
#endif
    out = lin_finish(acc);
    return true;
  }
  if (op == OP_INV) {
    out = fp_inv_wvm(ld_slot(slots, field(w0, w1, 1)));
    return true;
  }
  return false;
}

// Run a program for the pairing this warp owns.  prog: rounds x 32 lanes x 2 uint4.
#if defined(BN254_WVM_PROFILE) && defined(__CUDACC__)
__device__ unsigned long long wvm_prof[8];  // cycles and rounds per op class (debug build: profiles/r2/wvm_round_costs.json)
#endif
BN_HD void run_round(Fp* slots, const uint4& w0, const uint4& w1, int lane) {
#if defined(BN254_WVM_PROFILE) && defined(__CUDACC__)
  long long t0_ = clock64();
  unsigned cls_ = __reduce_max_sync(0xffffffffu, w0.x & 3u);
#endif
  unsigned h = w0.x & 0xFFFFu;
  unsigned nterms = (h & 3u) == OP_LIN ? (h >> 12) : 0u;
#if defined(__CUDACC__)
  unsigned nmax = warp_max(nterms);
#else
  unsigned nmax = 15;
#endif
  unsigned dst = 0;
  Fp out;
  bool st = exec_op(slots, w0, w1, nmax, dst, out);
  if (st) st_slot(slots, dst, out);
  round_sync();
#if defined(BN254_WVM_PROFILE) && defined(__CUDACC__)
  if (lane == 0 && blockIdx.x == 0 && threadIdx.x == 0) { wvm_prof[cls_] += (unsigned long long)(clock64() - t0_); wvm_prof[4 + cls_] += 1; }
#endif
  (void)lane;
}
BN_HD void run(Fp* slots, const uint4* __restrict__ prog, int rounds, int lane) {
  // The program words of round r + kAhead are requested in round r, into a ring of kAhead register pairs that is indexed
  // STATICALLY (the round loop is unrolled by kAhead, no register shifting).  kAhead = 1 is the measured best: a lone warp
  // is bound by its ~375 instructions per round at ~3.3 cycles each, not by the fetch of its program words
  // (long_scoreboard 4.8 % of its stall samples, profiles/r2/ncu_k_wvm_pair1_lone_warp_summary.txt); rings 2 / 3 / 4
  // rounds deep measure 1.34 / 1.41 / 1.44 ms for a 1-element Pair against 1.31 ms (profiles/r2/wvm_prefetch_ring_ab.jsonl).
  constexpr int kAhead = WVM_AHEAD;
  const uint4* p = prog + (size_t)lane * 2;
  uint4 q0[kAhead], q1[kAhead];
#pragma unroll
  for (int k = 0; k < kAhead; k++) {
    int rr = k < rounds ? k : rounds - 1;
    q0[k] = p[(size_t)rr * kLanes * 2]; q1[k] = p[(size_t)rr * kLanes * 2 + 1];
  }
  for (int r0 = 0; r0 < rounds; r0 += kAhead) {
#pragma unroll
    for (int k = 0; k < kAhead; k++) {
      const int r = r0 + k;
      if (r < rounds) {  // warp-uniform
        const uint4 w0 = q0[k], w1 = q1[k];
        const int rr = r + kAhead < rounds ? r + kAhead : rounds - 1;
        q0[k] = p[(size_t)rr * kLanes * 2]; q1[k] = p[(size_t)rr * kLanes * 2 + 1];
        run_round(slots, w0, w1, lane);
      }
    }
  }
}

} }  // namespace bn254::wvm
