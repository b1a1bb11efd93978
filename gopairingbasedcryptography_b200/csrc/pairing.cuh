// Optimal-ate pairing on BN254: Miller loop and final exponentiation.
// Replaces (reference side): gnark-crypto v0.19.0 ecc/bn254/pairing.go {Pair, PairingCheck,
// MillerLoop, FinalExponentiation}, called at access/tree/access_tree_node.go:106,110,
// cpabe/bsw07/bsw07_cpabe.go:184, ibe/waters05_ibe/waters05_ibe.go:214,259,262,
// bibe/afp25_bibe/afp25_bibe.go:395-403, signature/bls01_signature/bls_signature.go:81-84.
//
// Conventions: D-type twist, untwist (x',y') -> (x' w^2, y' w^3); a line through
// twist points with slope lam evaluated at P=(xP,yP) is  yP - lam xP w + (lam xT - yT) w^3, i.e. the
// sparse Fp12 with non-zero slots (c0.b0, c1.b0, c1.b1).  Lines are kept up to Fp2 factors (killed by
// the final exponentiation).  Final exponent d' = 2x0(6x0^2+3x0+1) (p^12-1)/r as in gnark.
#pragma once
#include "tower.cuh"

namespace bn254 {

struct G1Aff { Fp x, y; };
struct G2Aff { Fp2 x, y; };
struct G2Proj { Fp2 x, y, z; };
static constexpr int kLinesPerPoint = 65 + 21 + 2;  // lines per fixed G2 point (line tables, below)

BN_HD bool g1_is_inf(const G1Aff& p) { return fp_is_zero(p.x) && fp_is_zero(p.y); }
BN_HD bool g2_is_inf(const G2Aff& q) { return fp2_is_zero(q.x) && fp2_is_zero(q.y); }

// Staged G2 steps (see tower_staged.cuh): T is copied into the scratch once (one local-memory round trip), every
// Fp2 product then runs out of shared memory, the new T goes straight back to memory and the raw line
// (r0, r1, r2) is LEFT in sc[3..5], where apply_line_sc consumes it -- the coefficients never visit local memory.
//
// Tangent at T (homogeneous projective), T <- 2T.  With E = 3b'Z^2:
//   X3 = XY/2 (Y^2 - 3E), Y3 = ((Y^2+3E)/2)^2 - 3E^2, Z3 = 2Y^3Z
//   line (times a subfield factor) = (-2YZ) yP + (3X^2) xP w + (E - Y^2) w^3
BN_NOINLINE void g2_dbl_step_sc(G2Proj& T, Fp2* sc) {
  BN_SC_REBIND(sc)
  BN_CTA_SYNC();
  stage6(sc, reinterpret_cast<const Fp6&>(T));  // X, Y, Z
  fp2_sqr(sc[3], sc[1]);                        // B = Y^2
  fp2_sqr(sc[4], sc[2]);                        // C = Z^2
  sc[8] = fp2_add(sc[1], sc[2]);
  fp2_sqr(sc[8], sc[8]);
  sc[8] = fp2_sub2(sc[8], sc[3], sc[4]);        // H = (Y+Z)^2 - B - C
  fp2_sqr(sc[6], sc[0]);                        // J = X^2
  fp2_mul(sc[7], sc[0], sc[1]);
  sc[7] = fp2_half(sc[7]);                      // A = XY/2
  fp2_mul(sc[4], sc[4], TWIST_3B);              // E = 3b' C
  sc[0] = fp2_triple(sc[4]);                    // F = 3E
  sc[1] = fp2_sub(sc[3], sc[0]);
  fp2_mul(T.x, sc[7], sc[1]);                   // X3 = A (B - F)
  sc[1] = fp2_add_half(sc[3], sc[0]);           // G = (B + F)/2
  fp2_sqr(sc[1], sc[1]);
  fp2_sqr(sc[2], sc[4]);                        // E^2
  T.y = fp2_sub_triple(sc[1], sc[2]);           // Y3 = G^2 - 3E^2
  fp2_mul(T.z, sc[3], sc[8]);                   // Z3 = B H
  sc[5] = fp2_sub(sc[4], sc[3]);                // r2 = E - B
  sc[3] = fp2_neg(sc[8]);                       // r0 = -H
  sc[4] = fp2_triple(sc[6]);                    // r1 = 3J
}
// Chord through T and affine Q (negated when neg), T <- T + Q.  O = Y1 - y2 Z1, L = X1 - x2 Z1:
//   line = L yP - O xP w + (O x2 - L y2) w^3
BN_NOINLINE void g2_stage_tq(Fp2* sc, const G2Proj& T, const G2Aff& Q, bool neg) {
  Fp2 x = fp2_ld(T.x), y = fp2_ld(T.y), z = fp2_ld(T.z), qx = fp2_ld(Q.x), qy = fp2_ld(Q.y);
  if (neg) qy = fp2_neg_i(qy);
  fp2_st(sc[0], x); fp2_st(sc[1], y); fp2_st(sc[2], z); fp2_st(sc[6], qx); fp2_st(sc[7], qy);
}
BN_NOINLINE void g2_add_step_sc(G2Proj& T, const G2Aff& Q, bool neg, bool update, Fp2* sc) {
  BN_SC_REBIND(sc)
  BN_CTA_SYNC();
  g2_stage_tq(sc, T, Q, neg);
  fp2_mul_rsub(sc[8], sc[7], sc[2], sc[1]);     // O = Y - y2 Z
  fp2_mul_rsub(sc[3], sc[6], sc[2], sc[0]);     // L = X - x2 Z            (r0)
  fp2_mul(sc[5], sc[6], sc[8]);
  fp2_mul_rsub(sc[5], sc[3], sc[7], sc[5]);     // r2 = x2 O - L y2
  sc[4] = fp2_neg(sc[8]);                       // r1 = -O
  if (!update) return;
  fp2_sqr(sc[6], sc[8]);                        // C = O^2
  fp2_sqr(sc[7], sc[3]);                        // D = L^2
  fp2_mul(sc[6], sc[2], sc[6]);                 // F = Z C
  fp2_mul(sc[0], sc[0], sc[7]);                 // G = X D
  fp2_mul(sc[7], sc[3], sc[7]);                 // E = L D
  sc[6] = fp2_add_sub_dbl(sc[7], sc[6], sc[0]); // H = E + F - 2G
  fp2_mul(T.x, sc[3], sc[6]);                   // X3 = L H
  fp2_mul(sc[1], sc[1], sc[7]);                 // Y E
  sc[0] = fp2_sub(sc[0], sc[6]);
  fp2_mul(sc[0], sc[0], sc[8]);                 // (G - H) O
  T.y = fp2_sub(sc[0], sc[1]);
  fp2_mul(T.z, sc[7], sc[2]);                   // Z3 = E Z
}
// f *= line(P), raw line in sc[3..5]
BN_NOINLINE void apply_line_sc(Fp12& f, const G1Aff& P, Fp2* sc) { BN_SC_REBIND(sc) apply_line_staged(f, P.x, P.y, sc + 3, sc + 6, sc, nullptr); }
// f *= line(P), raw line (r0, r1, r2) in memory (line tables): fetched together with f.c1, one round trip
BN_NOINLINE void apply_line_mem(Fp12& f, const G1Aff& P, const Fp2* r, Fp2* sc) { BN_SC_REBIND(sc) apply_line_staged(f, P.x, P.y, sc + 3, sc + 6, sc, r); }
// f *= line_a(Pa) * line_b(Pb), both raw lines in memory (line tables): see tower_staged.cuh
BN_NOINLINE void apply_line_pair_mem(Fp12& f, const G1Aff& Pa, const Fp2* ra, const G1Aff& Pb, const Fp2* rb, Fp2* sc) {
  BN_SC_REBIND(sc)
  Fp12 y;
  line_pair_stage(sc, ra, Pa.x, Pa.y, rb, Pb.x, Pb.y);
  fp12_mul_034_034(y, sc);
  fp12_mul_by_01234(f, y);
}
// One schedule position s of the line-table Miller loop for a chunk of `cnt` table points (bit j of skip: pair j holds a
// point at infinity): the lines are taken two at a time, a left-over one alone.
BN_HD void miller_lines_step(Fp12& f, const G1Aff* p, const Fp2* table, int first, int cnt, unsigned skip, int s, Fp2* sc) {
  int pending = -1;
  for (int j = 0; j < cnt; j++) {
    if ((skip >> j) & 1u) continue;
    if (pending < 0) { pending = j; continue; }
    const Fp2* La = table + ((size_t)(first + pending) * kLinesPerPoint + s) * 3;
    const Fp2* Lb = table + ((size_t)(first + j) * kLinesPerPoint + s) * 3;
    apply_line_pair_mem(f, p[pending], La, p[j], Lb, sc);
    pending = -1;
  }
  if (pending >= 0) apply_line_mem(f, p[pending], table + ((size_t)(first + pending) * kLinesPerPoint + s) * 3, sc);
}
BN_HD void apply_line(Fp12& f, const G1Aff& P, const Fp2& r0, const Fp2& r1, const Fp2& r2) {
  BN_SCRATCH_DECL
  Fp2 r[3] = {r0, r1, r2};
  apply_line_mem(f, P, r, sc_);
}
// compatibility forms writing the line to memory (line-table precomputation)
BN_HD void g2_dbl_step(G2Proj& T, Fp2& r0, Fp2& r1, Fp2& r2) {
  BN_SCRATCH_DECL
  g2_dbl_step_sc(T, sc_);
  r0 = sc_[3]; r1 = sc_[4]; r2 = sc_[5];
}
BN_HD void g2_add_step(G2Proj& T, const G2Aff& Q, Fp2& r0, Fp2& r1, Fp2& r2, bool update) {
  BN_SCRATCH_DECL
  g2_add_step_sc(T, Q, false, update, sc_);
  r0 = sc_[3]; r1 = sc_[4]; r2 = sc_[5];
}

// One G2 step (kind 0: tangent; 1: chord with +-Q[j]; 2: chord with q1; 3: chord with q2, no update) for every live pair,
// lines applied to f two at a time: the first line of a couple is parked on the stack while the second pair's step uses
// the scratch, then both go through ONE multiplication of f (apply_line_pair_mem); a left-over line is applied alone.
template <int KC>
BN_HD void miller_lines_of_step(Fp12& f, const G1Aff* P, const G2Aff* Q, G2Proj* T, int k, unsigned skip, int kind, bool neg, Fp2* sc_) {
  int pending = -1, last = -1;
  for (int j = 0; j < k; j++) if (!((skip >> j) & 1u)) last = j;
  Fp2 park[3];
  for (int j = 0; j < k; j++) {
    if ((skip >> j) & 1u) continue;
    if (kind == 0) g2_dbl_step_sc(T[j], sc_);
    else if (kind == 1) g2_add_step_sc(T[j], Q[j], neg, true, sc_);
    else {
      G2Aff q;
      if (kind == 2) { fp2_mul(q.x, fp2_conj(Q[j].x), GAMMA1[2]); fp2_mul(q.y, fp2_conj(Q[j].y), GAMMA1[3]); }
      else { q.x = fp2_mul_fp(Q[j].x, GAMMA2[2]); q.y = Q[j].y; }  // -pi^2(Q): xi^((p^2-1)/2) = -1
      g2_add_step_sc(T[j], q, false, kind == 2, sc_);
    }
    if (pending < 0) {
      if (j == last) { apply_line_sc(f, P[j], sc_); continue; }  // nobody to pair with: straight from the scratch
      Fp2 a = fp2_ld(sc_[3]), b = fp2_ld(sc_[4]), c = fp2_ld(sc_[5]);
      fp2_st(park[0], a); fp2_st(park[1], b); fp2_st(park[2], c);
      pending = j;
      continue;
    }
    apply_line_pair_mem(f, P[pending], park, P[j], sc_ + 3, sc_);
    pending = -1;
  }
}
template <int KC>
BN_HD void miller_loop_t(Fp12& f, const G1Aff* P, const G2Aff* Q, G2Proj* T, int k_rt) {
  BN_SCRATCH_DECL
  const int k = KC > 0 ? KC : k_rt;
  fp12_set_one(f);
  unsigned skip = 0;  // bit j set: pair j contains the point at infinity
  for (int j = 0; j < k; j++) {
    if (g1_is_inf(P[j]) || g2_is_inf(Q[j])) skip |= 1u << j;
    T[j].x = Q[j].x; T[j].y = Q[j].y; T[j].z = fp2_one();
  }
  if (skip == (k >= 32 ? 0xffffffffu : ((1u << k) - 1u))) return;
  if (KC >= 1 && KC <= 3) {
    // small compile-time pair counts (k_pair, the 2-pair BLS check, 3-pair products): every line straight from the scratch.
    // Pairing the lines was measured on them too: k_pair -2.7 % (instruction footprint), the 2-pair check neutral in
    // time with twice the local-memory traffic (a parked line per step) -- so they keep the direct path.
    for (int i = ATE_NAF_LEN - 2; i >= 0; i--) {
      if (i != ATE_NAF_LEN - 2) fp12_sqr(f, f);
      int d = ATE_NAF[i];
#pragma unroll
      for (int j = 0; j < KC; j++) {
        if ((skip >> j) & 1u) continue;
        g2_dbl_step_sc(T[j], sc_);
#ifdef BN254_PAIR_ADD_LINES
        if (d) {  // tangent and chord of this step as ONE multiplication of f (the tangent waits on the stack)
          Fp2 park[3];
          { Fp2 a = fp2_ld(sc_[3]), b = fp2_ld(sc_[4]), c = fp2_ld(sc_[5]); fp2_st(park[0], a); fp2_st(park[1], b); fp2_st(park[2], c); }
          g2_add_step_sc(T[j], Q[j], d < 0, true, sc_);
          apply_line_pair_mem(f, P[j], park, P[j], sc_ + 3, sc_);
        } else {
          apply_line_sc(f, P[j], sc_);
        }
#else
        apply_line_sc(f, P[j], sc_);
        if (d) {
          g2_add_step_sc(T[j], Q[j], d < 0, true, sc_);
          apply_line_sc(f, P[j], sc_);
        }
#endif
      }
    }
#pragma unroll
    for (int j = 0; j < KC; j++) {
      if ((skip >> j) & 1u) continue;
      G2Aff q1, q2;
      fp2_mul(q1.x, fp2_conj(Q[j].x), GAMMA1[2]);
      fp2_mul(q1.y, fp2_conj(Q[j].y), GAMMA1[3]);
      q2.x = fp2_mul_fp(Q[j].x, GAMMA2[2]); q2.y = Q[j].y;  // -pi^2(Q): xi^((p^2-1)/2) = -1
      g2_add_step_sc(T[j], q1, false, true, sc_);
      apply_line_sc(f, P[j], sc_);
      g2_add_step_sc(T[j], q2, false, false, sc_);
      apply_line_sc(f, P[j], sc_);
    }
    return;
  }
  for (int i = ATE_NAF_LEN - 2; i >= 0; i--) {
    if (i != ATE_NAF_LEN - 2) fp12_sqr(f, f);
    int d = ATE_NAF[i];
    miller_lines_of_step<KC>(f, P, Q, T, k, skip, 0, false, sc_);
    if (d) miller_lines_of_step<KC>(f, P, Q, T, k, skip, 1, d < 0, sc_);
  }
  miller_lines_of_step<KC>(f, P, Q, T, k, skip, 2, false, sc_);
  miller_lines_of_step<KC>(f, P, Q, T, k, skip, 3, false, sc_);
}
BN_HD void miller_loop(Fp12& f, const G1Aff* P, const G2Aff* Q, G2Proj* T, int k) {
  if (k == 1) miller_loop_t<1>(f, P, Q, T, 1);
  else miller_loop_t<0>(f, P, Q, T, k);
}


// ---- precomputed G2 line tables (north_star: fixed public-parameter / user-key G2 points) -----------------
// The (r0, r1, r2) coefficients of every line of the Miller schedule depend on Q only.  For a fixed Q they are
// computed once (g2_precompute_lines) and every later pairing against Q costs per line: 2 Fp2 x Fp products
// (evaluation at P) + one sparse 034 multiply -- no G2 arithmetic, no T state.
// Order: for i = 64..0 { tangent line; chord line if NAF digit != 0 }, then the two Frobenius lines.
BN_HD void g2_precompute_lines(const G2Aff& Q, Fp2* out /* kLinesPerPoint x 3 */) {
  G2Proj T; T.x = Q.x; T.y = Q.y; T.z = fp2_one();
  G2Aff qn; qn.x = Q.x; qn.y = fp2_neg(Q.y);
  int s = 0;
  for (int i = ATE_NAF_LEN - 2; i >= 0; i--) {
    g2_dbl_step(T, out[3 * s], out[3 * s + 1], out[3 * s + 2]); s++;
    int d = ATE_NAF[i];
    if (d) { g2_add_step(T, d > 0 ? Q : qn, out[3 * s], out[3 * s + 1], out[3 * s + 2], true); s++; }
  }
  G2Aff q1, q2;
  fp2_mul(q1.x, fp2_conj(Q.x), GAMMA1[2]);
  fp2_mul(q1.y, fp2_conj(Q.y), GAMMA1[3]);
  q2.x = fp2_mul_fp(Q.x, GAMMA2[2]); q2.y = Q.y;
  g2_add_step(T, q1, out[3 * s], out[3 * s + 1], out[3 * s + 2], true); s++;
  g2_add_step(T, q2, out[3 * s], out[3 * s + 1], out[3 * s + 2], false); s++;
}

// z^(d'), d' = 2x0(6x0^2+3x0+1)(p^12-1)/r.  Returns 1 early when the easy part is 1 (gnark behaviour).
// z == 0 is mapped to 0 by the inversion convention inv(0) = 0.
BN_NOINLINE void final_exp(Fp12& out, const Fp12& in) {
  Fp12 f, t0, t1, t2, t3, t4;
  fp12_conj(t0, in); fp12_inv(f, in); fp12_mul(t0, t0, f);
  fp12_frob(f, t0, 2); fp12_mul(f, f, t0);
  // gnark returns early when the easy part is 1.  The hard part maps 1 to 1, so under CTA lockstep (where a
  // data-dependent exit would leave the other warps waiting at a barrier) the thread simply computes on.
  if (!cta_lockstep_on() && fp12_is_one(f)) { out = f; return; }
  fp12_expt(t0, f); fp12_conj(t0, t0); fp12_cyclo_sqr(t0, t0);
  fp12_cyclo_sqr(t1, t0); fp12_mul(t1, t0, t1);
  fp12_expt(t2, t1); fp12_conj(t2, t2);
  fp12_conj(t3, t1); fp12_mul(t1, t2, t3);
  fp12_cyclo_sqr(t3, t2); fp12_expt(t4, t3); fp12_mul(t4, t1, t4);
  fp12_mul(t3, t0, t4); fp12_mul(t0, t2, t4); fp12_mul(t0, f, t0);
  fp12_frob(t2, t3, 1); fp12_mul(t0, t2, t0);
  fp12_frob(t2, t4, 2); fp12_mul(t0, t2, t0);
  fp12_conj(t2, f); fp12_mul(t2, t2, t3); fp12_frob(t2, t2, 3); fp12_mul(t0, t2, t0);
  out = t0;
}

}  // namespace bn254
