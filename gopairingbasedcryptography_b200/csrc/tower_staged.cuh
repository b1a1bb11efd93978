// Staged tower routines: the hot Fp6/Fp12 composites (included from tower.cuh INSIDE namespace bn254).
//
// Why: in the one-thread-per-pairing kernels every Fp12/Fp6 value lives on the per-thread local-memory stack.
// With 227 KB of the SM given to the shared-memory scratch the L1 keeps ~75 B per thread, so every local
// operand costs an L2 round trip (ncu: long_scoreboard 1.4 cycles per issued instruction, IADD3/CALL samples
// 45-80 % memory waits).  The unstaged routines paid one such round trip per Fp2 leaf (~40 per Fp12 product).
// Here each composite routine
//   1. copies its Fp6 operands into the thread's 9-slot shared-memory scratch with ALL loads in flight at once
//      (one exposed round trip per 6 Fp2),
//   2. runs every Fp2 product out of the scratch, with the Karatsuba pre-additions and the recombination fused
//      into the multiply leaf (fp2_cross: no intermediate stores, 7 calls per Fp6 product instead of 21),
//   3. writes results straight to their destination; operands needed by a recombination are fetched BEFORE the
//      leaf's three Montgomery products so their latency hides behind ~1600 multiply-pipe cycles.
// The arithmetic (which Fp products, which reductions) is unchanged, so results stay bit-identical.
//
// Replaces (reference side): gnark-crypto v0.19.0 ecc/bn254/internal/fptower/{e6,e12,e12_pairing}.go
// (E6.Mul, E12.Mul, E12.Square, E12.CyclotomicSquare, E12.MulBy034), reached from bn254.Pair / GT.Mul / GT.Exp.
#pragma once

// ---- small by-value leaves used between scratch slots ------------------------------------------------------
BN_LEAF Fp2 fp2_sub2(const Fp2& a, const Fp2& b, const Fp2& c) { return fp2_sub_i(fp2_sub_i(fp2_ld(a), fp2_ld(b)), fp2_ld(c)); }
BN_LEAF Fp2 fp2_triple(const Fp2& a) { Fp2 v = fp2_ld(a); return fp2_add_i(fp2_dbl_i(v), v); }
BN_LEAF Fp2 fp2_sub_triple(const Fp2& a, const Fp2& b) { Fp2 v = fp2_ld(b); return fp2_sub_i(fp2_ld(a), fp2_add_i(fp2_dbl_i(v), v)); }
BN_LEAF Fp2 fp2_add_sub_dbl(const Fp2& a, const Fp2& b, const Fp2& c) { return fp2_sub_i(fp2_add_i(fp2_ld(a), fp2_ld(b)), fp2_dbl_i(fp2_ld(c))); }
BN_LEAF Fp2 fp2_add_half(const Fp2& a, const Fp2& b) { return fp2_half_i(fp2_add_i(fp2_ld(a), fp2_ld(b))); }

// xi-multiplication with operand and result in registers: ONE copy of its ~250 instructions serves every fused
// leaf below (inlining it made fp2_cross / fp2_cyc 19 KB each and the kernel instruction-cache bound).
BN_NOINLINE Fp2 fp2_mul_xi_bv(Fp2 a) { return fp2_mul_xi_i(a); }

// z = c - a*b
BN_NOINLINE void fp2_mul_rsub(Fp2& z, const Fp2& a, const Fp2& b, const Fp2& c) {
  Fp2 C = fp2_ld(c);
  fp2_st(z, fp2_sub_i(C, FP2_MUL(fp2_ld(a), fp2_ld(b))));
}

// Karatsuba cross term with the recombination fused in:
//   t = (xa + xb) * (ya [+ yb]) - va [- vb]
//   mode 0: r = xi*t + vc    mode 1: r = t + xi*vc    mode 2: r = t + vc    mode 3: r = t
//   z = r [- e0] [- e1]
// Everything that does not depend on the product is folded into ONE correction K before the multiplication, so the
// dependent chain after the product is a single addition (modes 1-3: z = prod + K, K = c - va - vb - e0 - e1 with
// c = xi*vc, vc or 0) or subtract / xi / add (mode 0: z = xi*(prod - S) + K, S = va + vb, K = vc - e0 - e1 -- one xi
// per call in every mode).  e0/e1 usually live in local memory; their latency overlaps the pre-additions.
enum { kCrossXiT = 0, kCrossXiV = 1, kCrossPlain = 2, kCrossNone = 3 };
BN_NOINLINE void fp2_cross(Fp2& z, const Fp2& xa, const Fp2& xb, const Fp2& ya, const Fp2* yb, const Fp2& va, const Fp2* vb,
                           const Fp2* vc, int mode, const Fp2* e0, const Fp2* e1) {
  Fp2 S = fp2_ld(va);
  if (vb) S = fp2_add_i(S, fp2_ld(*vb));
  Fp2 K;
  if (mode == kCrossXiT) {
    K = fp2_ld(*vc);
    if (e0) K = fp2_sub_i(K, fp2_ld(*e0));
    if (e1) K = fp2_sub_i(K, fp2_ld(*e1));
  } else {
    if (e0) S = fp2_add_i(S, fp2_ld(*e0));
    if (e1) S = fp2_add_i(S, fp2_ld(*e1));
    if (mode == kCrossNone) K = fp2_neg_i(S);
    else {
      Fp2 c = fp2_ld(*vc);
      if (mode == kCrossXiV) c = fp2_mul_xi_bv(c);
      K = fp2_sub_i(c, S);
    }
  }
  Fp2 x = fp2_add_i(fp2_ld(xa), fp2_ld(xb));
  Fp2 y = fp2_ld(ya);
  if (yb) y = fp2_add_i(y, fp2_ld(*yb));
  Fp2 t = FP2_MUL(x, y);
  if (mode == kCrossXiT) t = fp2_mul_xi_bv(fp2_sub_i(t, S));
  fp2_st(z, fp2_add_i(t, K));
}

// ---- staging: generic memory -> scratch slots, every load issued before the first store ---------------------
BN_NOINLINE void stage6(Fp2* dst, const Fp6& x) {
  Fp2 a = fp2_ld(x.b0), b = fp2_ld(x.b1), c = fp2_ld(x.b2);
  fp2_st(dst[0], a); fp2_st(dst[1], b); fp2_st(dst[2], c);
}
BN_NOINLINE void stage6x2(Fp2* dst, const Fp6& x, const Fp6& y) {
  Fp2 a = fp2_ld(x.b0), b = fp2_ld(x.b1), c = fp2_ld(x.b2), d = fp2_ld(y.b0), e = fp2_ld(y.b1), f = fp2_ld(y.b2);
  fp2_st(dst[0], a); fp2_st(dst[1], b); fp2_st(dst[2], c); fp2_st(dst[3], d); fp2_st(dst[4], e); fp2_st(dst[5], f);
}
// dst[0..2] = x + y
BN_NOINLINE void stage6_sum(Fp2* dst, const Fp6& x, const Fp6& y) {
  Fp2 a = fp2_ld(x.b0), b = fp2_ld(x.b1), c = fp2_ld(x.b2), d = fp2_ld(y.b0), e = fp2_ld(y.b1), f = fp2_ld(y.b2);
  fp2_st(dst[0], fp2_add_i(a, d)); fp2_st(dst[1], fp2_add_i(b, e)); fp2_st(dst[2], fp2_add_i(c, f));
}

// ---- Fp6 -----------------------------------------------------------------------------------------------------
// x in sc[0..2], y in sc[3..5]; sc[6..8] receive the diagonal products.  z = x*y [- e0 - e1] (z: any memory;
// the operands were staged, so z may alias whatever they were copied from).
BN_HD void fp6_mul_staged(Fp6& z, Fp2* sc, const Fp6* e0, const Fp6* e1) {
  BN_CTA_SYNC();
  fp2_mul(sc[6], sc[0], sc[3]);
  fp2_mul(sc[7], sc[1], sc[4]);
  fp2_mul(sc[8], sc[2], sc[5]);
  fp2_cross(z.b0, sc[1], sc[2], sc[4], &sc[5], sc[7], &sc[8], &sc[6], kCrossXiT, e0 ? &e0->b0 : nullptr, e1 ? &e1->b0 : nullptr);
  fp2_cross(z.b1, sc[0], sc[1], sc[3], &sc[4], sc[6], &sc[7], &sc[8], kCrossXiV, e0 ? &e0->b1 : nullptr, e1 ? &e1->b1 : nullptr);
  fp2_cross(z.b2, sc[0], sc[2], sc[3], &sc[5], sc[6], &sc[8], &sc[7], kCrossPlain, e0 ? &e0->b2 : nullptr, e1 ? &e1->b2 : nullptr);
}
BN_NOINLINE void fp6_mul(Fp6& z, const Fp6& x, const Fp6& y) {
  BN_SCRATCH_DECL
  stage6x2(sc_, x, y);
  fp6_mul_staged(z, sc_, nullptr, nullptr);
}
// x in X[0..2]; z = x * (c0 + c1 v) [- e0 - e1]; a, b: two free slots
BN_HD void fp6_mul_01_staged(Fp6& z, const Fp2* X, const Fp2& c0, const Fp2& c1, Fp2& a, Fp2& b, const Fp6* e0, const Fp6* e1) {
  BN_CTA_SYNC();
  fp2_mul(a, X[0], c0);
  fp2_mul(b, X[1], c1);
  fp2_cross(z.b0, X[1], X[2], c1, nullptr, b, nullptr, &a, kCrossXiT, e0 ? &e0->b0 : nullptr, e1 ? &e1->b0 : nullptr);
  fp2_cross(z.b2, X[0], X[2], c0, nullptr, a, nullptr, &b, kCrossPlain, e0 ? &e0->b2 : nullptr, e1 ? &e1->b2 : nullptr);
  fp2_cross(z.b1, X[0], X[1], c0, &c1, a, &b, nullptr, kCrossNone, e0 ? &e0->b1 : nullptr, e1 ? &e1->b1 : nullptr);
}
// z = a + v*b
BN_NOINLINE void fp6_add_mul_v(Fp6& z, const Fp6& a, const Fp6& b) {
  Fp2 a0 = fp2_ld(a.b0), a1 = fp2_ld(a.b1), a2 = fp2_ld(a.b2), b0 = fp2_ld(b.b0), b1 = fp2_ld(b.b1), b2 = fp2_ld(b.b2);
  fp2_st(z.b0, fp2_add_i(a0, fp2_mul_xi_bv(b2)));
  fp2_st(z.b1, fp2_add_i(a1, b0));
  fp2_st(z.b2, fp2_add_i(a2, b1));
}

// ---- Fp12 ----------------------------------------------------------------------------------------------------
// Karatsuba over Fp6: a = x0 y0, b = x1 y1, z1 = (x0+x1)(y0+y1) - a - b, z0 = a + v b.  z may alias x or y.
BN_NOINLINE void fp12_mul(Fp12& z, const Fp12& x, const Fp12& y) {
  BN_SCRATCH_DECL
  Fp6 a, b;
  stage6x2(sc_, x.c0, y.c0);
  fp6_mul_staged(a, sc_, nullptr, nullptr);
  stage6x2(sc_, x.c1, y.c1);
  fp6_mul_staged(b, sc_, nullptr, nullptr);
  stage6_sum(sc_, x.c0, x.c1);
  stage6_sum(sc_ + 3, y.c0, y.c1);
  fp6_mul_staged(z.c1, sc_, &a, &b);
  fp6_add_mul_v(z.c0, a, b);
}
// x * conj(y), y = (y0, y1) -> (y0, -y1):  a = x0 y0, b = x1 y1, z1 = (x0+x1)(y0-y1) - a + b, z0 = a - v b
BN_NOINLINE void stage6_diff(Fp2* dst, const Fp6& x, const Fp6& y) {
  Fp2 a = fp2_ld(x.b0), b = fp2_ld(x.b1), c = fp2_ld(x.b2), d = fp2_ld(y.b0), e = fp2_ld(y.b1), f = fp2_ld(y.b2);
  fp2_st(dst[0], fp2_sub_i(a, d)); fp2_st(dst[1], fp2_sub_i(b, e)); fp2_st(dst[2], fp2_sub_i(c, f));
}
BN_NOINLINE void fp12_mul_conj_tail(Fp12& z, const Fp6& a, const Fp6& b, const Fp6& s) {
  // z1 = s - a + b ; z0 = a - v b
  Fp2 a0 = fp2_ld(a.b0), a1 = fp2_ld(a.b1), a2 = fp2_ld(a.b2), b0 = fp2_ld(b.b0), b1 = fp2_ld(b.b1), b2 = fp2_ld(b.b2);
  Fp2 s0 = fp2_ld(s.b0), s1 = fp2_ld(s.b1), s2 = fp2_ld(s.b2);
  fp2_st(z.c1.b0, fp2_add_i(fp2_sub_i(s0, a0), b0));
  fp2_st(z.c1.b1, fp2_add_i(fp2_sub_i(s1, a1), b1));
  fp2_st(z.c1.b2, fp2_add_i(fp2_sub_i(s2, a2), b2));
  fp2_st(z.c0.b0, fp2_sub_i(a0, fp2_mul_xi_bv(b2)));
  fp2_st(z.c0.b1, fp2_sub_i(a1, b0));
  fp2_st(z.c0.b2, fp2_sub_i(a2, b1));
}
BN_NOINLINE void fp12_mul_conj(Fp12& z, const Fp12& x, const Fp12& y) {
  BN_SCRATCH_DECL
  Fp6 a, b, s;
  stage6x2(sc_, x.c0, y.c0);
  fp6_mul_staged(a, sc_, nullptr, nullptr);
  stage6x2(sc_, x.c1, y.c1);
  fp6_mul_staged(b, sc_, nullptr, nullptr);
  stage6_sum(sc_, x.c0, x.c1);
  stage6_diff(sc_ + 3, y.c0, y.c1);
  fp6_mul_staged(s, sc_, nullptr, nullptr);
  fp12_mul_conj_tail(z, a, b, s);
}
// complex squaring: m = x0 x1, s = (x0 + x1)(x0 + v x1); z0 = s - m - v m, z1 = 2 m
BN_NOINLINE void fp12_sqr_prep(Fp2* sc) {  // (x0, x1) in sc[0..5] -> (x0 + x1, x0 + v x1), in place
  Fp2 a0 = fp2_ld(sc[0]), a1 = fp2_ld(sc[1]), a2 = fp2_ld(sc[2]), b0 = fp2_ld(sc[3]), b1 = fp2_ld(sc[4]), b2 = fp2_ld(sc[5]);
  fp2_st(sc[0], fp2_add_i(a0, b0)); fp2_st(sc[1], fp2_add_i(a1, b1)); fp2_st(sc[2], fp2_add_i(a2, b2));
  fp2_st(sc[3], fp2_add_i(a0, fp2_mul_xi_bv(b2))); fp2_st(sc[4], fp2_add_i(a1, b0)); fp2_st(sc[5], fp2_add_i(a2, b1));
}
BN_NOINLINE void fp12_sqr_tail(Fp12& z, const Fp6& s, const Fp6& m) {
  Fp2 s0 = fp2_ld(s.b0), s1 = fp2_ld(s.b1), s2 = fp2_ld(s.b2), m0 = fp2_ld(m.b0), m1 = fp2_ld(m.b1), m2 = fp2_ld(m.b2);
  fp2_st(z.c0.b0, fp2_sub_i(fp2_sub_i(s0, m0), fp2_mul_xi_bv(m2)));
  fp2_st(z.c0.b1, fp2_sub_i(fp2_sub_i(s1, m1), m0));
  fp2_st(z.c0.b2, fp2_sub_i(fp2_sub_i(s2, m2), m1));
  fp2_st(z.c1.b0, fp2_dbl_i(m0)); fp2_st(z.c1.b1, fp2_dbl_i(m1)); fp2_st(z.c1.b2, fp2_dbl_i(m2));
}
BN_NOINLINE void fp12_sqr(Fp12& z, const Fp12& x) {
  BN_SCRATCH_DECL
  Fp6 m, s;
  stage6x2(sc_, x.c0, x.c1);
  fp6_mul_staged(m, sc_, nullptr, nullptr);
  fp12_sqr_prep(sc_);
  fp6_mul_staged(s, sc_, nullptr, nullptr);
  fp12_sqr_tail(z, s, m);
}

// ---- Granger-Scott cyclotomic squaring -----------------------------------------------------------------------
// Fp12 = Fp4[w]/(w^3 - s); in the w-basis g0=c0.b0 g1=c1.b0 g2=c0.b1 g3=c1.b1 g4=c0.b2 g5=c1.b2 and with
// (r0, r1) = (a^2 + xi b^2, 2ab) for the pairs (g0,g3), (g1,g4), (g2,g5):
//   g0' = 3 r0(g0,g3) - 2 g0   g3' = 3 r1(g0,g3) + 2 g3
//   g2' = 3 r0(g1,g4) - 2 g2   g5' = 3 r1(g1,g4) + 2 g5
//   g4' = 3 r0(g2,g5) - 2 g4   g1' = 3 xi r1(g2,g5) + 2 g1
// One leaf per pair: 3 Fp2 squarings and both outputs, everything in registers.
BN_NOINLINE void fp2_cyc(Fp2& zm, Fp2& zp, const Fp2& a, const Fp2& b, const Fp2& gm, const Fp2& gp, int xi) {
#ifdef BN254_CYC_LATE_LOADS
  // 128-register builds: the linear terms are fetched after the squarings (two 16-register values less across six calls)
  Fp2 A = fp2_ld(a), B = fp2_ld(b);
  Fp2 S = FP2_SQR(fp2_add_i(A, B));
  A = FP2_SQR(A);
  B = FP2_SQR(B);
  Fp2 r1 = fp2_sub_i(fp2_sub_i(S, A), B);
  Fp2 r0 = fp2_add_i(A, fp2_mul_xi_bv(B));
  if (xi) r1 = fp2_mul_xi_bv(r1);
  Fp2 Gm = fp2_ld(gm), Gp = fp2_ld(gp);
#else
  Fp2 A = fp2_ld(a), B = fp2_ld(b), Gm = fp2_ld(gm), Gp = fp2_ld(gp);
  Fp2 S = FP2_SQR(fp2_add_i(A, B));
  A = FP2_SQR(A);
  B = FP2_SQR(B);
  Fp2 r1 = fp2_sub_i(fp2_sub_i(S, A), B);
  Fp2 r0 = fp2_add_i(A, fp2_mul_xi_bv(B));
  if (xi) r1 = fp2_mul_xi_bv(r1);
#endif
  fp2_st(zm, fp2_add_i(fp2_dbl_i(fp2_sub_i(r0, Gm)), r0));
  fp2_st(zp, fp2_add_i(fp2_dbl_i(fp2_add_i(r1, Gp)), r1));
}
// Slot map while an element is staged: memory order, slot j = j-th Fp2 of the Fp12
//   sc[0]=g0 sc[1]=g2 sc[2]=g4 sc[3]=g1 sc[4]=g3 sc[5]=g5
BN_NOINLINE void fp12_cyclo_sqr(Fp12& z, const Fp12& x) {
  BN_SCRATCH_DECL
  stage6x2(sc_, x.c0, x.c1);
  fp2_cyc(z.c0.b0, z.c1.b1, sc_[0], sc_[4], sc_[0], sc_[4], 0);
  fp2_cyc(z.c0.b1, z.c1.b2, sc_[3], sc_[2], sc_[1], sc_[5], 0);
  fp2_cyc(z.c0.b2, z.c1.b0, sc_[1], sc_[5], sc_[2], sc_[3], 1);
}
// n >= 1 successive squarings with the element resident in the scratch: one staging round trip per run.
// Pair 2 needs g2, g5 as linear terms and pair 3 squares them, so g2', g5' go to two spare slots and the
// (g2, g5) slots alternate between {1, 5} and {6, 7}.
BN_NOINLINE void fp12_cyclo_sqr_n(Fp12& z, const Fp12& x, int n) {
  BN_SCRATCH_DECL
  stage6x2(sc_, x.c0, x.c1);
  int i2 = 1, i5 = 5, f2 = 6, f5 = 7;
  for (int it = 0; it < n - 1; it++) {
    BN_CTA_SYNC();
    fp2_cyc(sc_[0], sc_[4], sc_[0], sc_[4], sc_[0], sc_[4], 0);
    fp2_cyc(sc_[f2], sc_[f5], sc_[3], sc_[2], sc_[i2], sc_[i5], 0);
    fp2_cyc(sc_[2], sc_[3], sc_[i2], sc_[i5], sc_[2], sc_[3], 1);
    int t = i2; i2 = f2; f2 = t;
    t = i5; i5 = f5; f5 = t;
  }
  BN_CTA_SYNC();
  fp2_cyc(z.c0.b0, z.c1.b1, sc_[0], sc_[4], sc_[0], sc_[4], 0);
  fp2_cyc(z.c0.b1, z.c1.b2, sc_[3], sc_[2], sc_[i2], sc_[i5], 0);
  fp2_cyc(z.c0.b2, z.c1.b0, sc_[i2], sc_[i5], sc_[2], sc_[3], 1);
}

// ---- sparse "034" line multiply ------------------------------------------------------------------------------
// f *= l0 + l1 w + l3 w^3 with f = (X, Y) over Fp6:  a = X l0, b = Y (l1 + l3 v), s = (X+Y)((l0+l1) + l3 v);
// f1 = s - a - b, f0 = a + v b.  13 Fp2 products.
// The raw line coefficients r0, r1, r2 are expected in three scratch slots L[0..2] (left there by the G2 step
// or copied from a line table); Y[0..2] and W[0..2] are the other six slots.
BN_NOINLINE void line_stage_y(Fp2* L, Fp2* Y, const Fp6& fy, const Fp& px, const Fp& py, const Fp2* raw) {
  // Y <- f.c1; L0 <- r0 * yP; L1 <- r1 * xP; L2 <- r2.  raw == nullptr: (r0, r1, r2) already sit in L (left by the
  // G2 step); otherwise they are read from memory (line table) in the same batch of loads as f.c1.
  Fp2 y0 = fp2_ld(fy.b0), y1 = fp2_ld(fy.b1), y2 = fp2_ld(fy.b2);
  Fp x = fp_ld(px), y = fp_ld(py);
  const Fp2* src = raw ? raw : L;
  Fp2 r0 = fp2_ld(src[0]), r1 = fp2_ld(src[1]);
  if (raw) { Fp2 r2 = fp2_ld(raw[2]); fp2_st(L[2], r2); }
  fp2_st(Y[0], y0); fp2_st(Y[1], y1); fp2_st(Y[2], y2);
  fp2_st(L[0], fp2_mul_fp_i(r0, y));
  fp2_st(L[1], fp2_mul_fp_i(r1, x));
}
BN_NOINLINE void line_sum_xy(Fp2* L, Fp2* Y, const Fp2* X) {  // Y += X ; L0 += L1
  Fp2 y0 = fp2_ld(Y[0]), y1 = fp2_ld(Y[1]), y2 = fp2_ld(Y[2]), x0 = fp2_ld(X[0]), x1 = fp2_ld(X[1]), x2 = fp2_ld(X[2]);
  Fp2 l0 = fp2_ld(L[0]), l1 = fp2_ld(L[1]);
  fp2_st(Y[0], fp2_add_i(y0, x0)); fp2_st(Y[1], fp2_add_i(y1, x1)); fp2_st(Y[2], fp2_add_i(y2, x2));
  fp2_st(L[0], fp2_add_i(l0, l1));
}
BN_HD void apply_line_staged(Fp12& f, const Fp& px, const Fp& py, Fp2* L, Fp2* Y, Fp2* W, const Fp2* raw) {
  Fp6 a, b;
  line_stage_y(L, Y, f.c1, px, py, raw);
  fp6_mul_01_staged(b, Y, L[1], L[2], W[0], W[1], nullptr, nullptr);
  stage6(W, f.c0);
  fp2_mul(a.b0, W[0], L[0]); fp2_mul(a.b1, W[1], L[0]); fp2_mul(a.b2, W[2], L[0]);
  line_sum_xy(L, Y, W);
  fp6_mul_01_staged(f.c1, Y, L[0], L[2], W[0], W[1], &a, &b);
  fp6_add_mul_v(f.c0, a, b);
}
// stand-alone form (line coefficients in memory): l0, l1, l3 already evaluated at P
BN_NOINLINE void fp12_mul_034(Fp12& z, const Fp2& l0, const Fp2& l1, const Fp2& l3) {
  BN_SCRATCH_DECL
  Fp2 *L = sc_ + 3, *Y = sc_ + 6, *W = sc_;
  Fp6 a, b;
  {
    Fp2 v0 = fp2_ld(l0), v1 = fp2_ld(l1), v3 = fp2_ld(l3);
    fp2_st(L[0], v0); fp2_st(L[1], v1); fp2_st(L[2], v3);
  }
  stage6(Y, z.c1);
  fp6_mul_01_staged(b, Y, L[1], L[2], W[0], W[1], nullptr, nullptr);
  stage6(W, z.c0);
  fp2_mul(a.b0, W[0], L[0]); fp2_mul(a.b1, W[1], L[0]); fp2_mul(a.b2, W[2], L[0]);
  line_sum_xy(L, Y, W);
  fp6_mul_01_staged(z.c1, Y, L[0], L[2], W[0], W[1], &a, &b);
  fp6_add_mul_v(z.c0, a, b);
}

// ---- two lines of one Miller step at once ----------------------------------------------------------------------------
// f *= l(P) * l'(P'):  the two sparse factors are multiplied first -- 034 x 034 has five non-zero coefficients, six Fp2
// products -- and f takes ONE multiplication by that 01234-sparse element (17 Fp2 products: 6 + 5 + 6) instead of two
// 034 multiplications (2 x 13); f is read and written once instead of twice.  Exact field arithmetic: same bytes as the
// two single-line multiplications in either order.
//   A = (a0, a3, a4) = (r0 yP, r1 xP, r2) on (c0.b0, c1.b0, c1.b1), B likewise:
//   y0 = a0 b0 + xi a4 b4   y1 = a3 b3   y2 = a3 b4 + a4 b3   y3 = a0 b3 + a3 b0   y4 = a0 b4 + a4 b0   (y5 = 0)
BN_LEAF Fp2 fp2_add_mul_xi(const Fp2& a, const Fp2& b) { return fp2_add_i(fp2_ld(a), fp2_mul_xi_bv(fp2_ld(b))); }
BN_NOINLINE void line_pair_stage(Fp2* sc, const Fp2* ra, const Fp& ax, const Fp& ay, const Fp2* rb, const Fp& bx, const Fp& by) {
  Fp2 a0 = fp2_ld(ra[0]), a1 = fp2_ld(ra[1]), a2 = fp2_ld(ra[2]), b0 = fp2_ld(rb[0]), b1 = fp2_ld(rb[1]), b2 = fp2_ld(rb[2]);
  Fp pax = fp_ld(ax), pay = fp_ld(ay), pbx = fp_ld(bx), pby = fp_ld(by);
  fp2_st(sc[2], a2); fp2_st(sc[5], b2);
  fp2_st(sc[0], fp2_mul_fp_i(a0, pay));
  fp2_st(sc[1], fp2_mul_fp_i(a1, pax));
  fp2_st(sc[3], fp2_mul_fp_i(b0, pby));
  fp2_st(sc[4], fp2_mul_fp_i(b1, pbx));
}
BN_NOINLINE void fp12_mul_034_034(Fp12& y, Fp2* sc) {  // operands in sc[0..5] (line_pair_stage), sc[6..8] temporaries
  BN_CTA_SYNC();
  fp2_mul(sc[6], sc[0], sc[3]);  // a0 b0
  fp2_mul(sc[7], sc[1], sc[4]);  // a3 b3
  fp2_mul(sc[8], sc[2], sc[5]);  // a4 b4
  fp2_cross(y.c0.b2, sc[1], sc[2], sc[4], &sc[5], sc[7], &sc[8], nullptr, kCrossNone, nullptr, nullptr);
  fp2_cross(y.c1.b0, sc[0], sc[1], sc[3], &sc[4], sc[6], &sc[7], nullptr, kCrossNone, nullptr, nullptr);
  fp2_cross(y.c1.b1, sc[0], sc[2], sc[3], &sc[5], sc[6], &sc[8], nullptr, kCrossNone, nullptr, nullptr);
  y.c0.b0 = fp2_add_mul_xi(sc[6], sc[8]);
  { Fp2 t = fp2_ld(sc[7]); fp2_st(y.c0.b1, t); }
  y.c1.b2 = fp2_zero();
}
BN_NOINLINE void stage_y01234(Fp2* dst, const Fp12& y, bool sum) {  // sum: (y0 + y3, y1 + y4, y2); else (y3, y4)
  if (sum) {
    Fp2 a = fp2_ld(y.c0.b0), b = fp2_ld(y.c0.b1), c = fp2_ld(y.c0.b2), d = fp2_ld(y.c1.b0), e = fp2_ld(y.c1.b1);
    fp2_st(dst[0], fp2_add_i(a, d)); fp2_st(dst[1], fp2_add_i(b, e)); fp2_st(dst[2], c);
  } else {
    Fp2 d = fp2_ld(y.c1.b0), e = fp2_ld(y.c1.b1);
    fp2_st(dst[0], d); fp2_st(dst[1], e);
  }
}
BN_NOINLINE void fp12_mul_by_01234(Fp12& f, const Fp12& y) {  // y.c1.b2 == 0
  BN_SCRATCH_DECL
  Fp6 a, b;
  stage6x2(sc_, f.c0, y.c0);
  fp6_mul_staged(a, sc_, nullptr, nullptr);
  stage6(sc_, f.c1);
  stage_y01234(sc_ + 3, y, false);
  fp6_mul_01_staged(b, sc_, sc_[3], sc_[4], sc_[5], sc_[6], nullptr, nullptr);
  stage6_sum(sc_, f.c0, f.c1);
  stage_y01234(sc_ + 3, y, true);
  fp6_mul_staged(f.c1, sc_, &a, &b);
  fp6_add_mul_v(f.c0, a, b);
}
