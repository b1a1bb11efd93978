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
BN_NOINLINE void fp_mul_ool(Fp& z, const Fp& a, const Fp& b) { z = fp_mul(a, b); }
BN_HD Fp f_mul(const Fp& a, const Fp& b) { Fp z; fp_mul_ool(z, a, b); return z; }
BN_HD Fp f_sqr(const Fp& a) { Fp z; fp_mul_ool(z, a, a); return z; }
BN_HD Fp f_inv(const Fp& a) { Fp z; fp_inv_ool(z, a); return z; }
BN_HD bool f_is_zero(const Fp& a) { return fp_is_zero(a); }
BN_HD void f_set_one(Fp& a) { a = fp_one(); }
BN_HD void f_set_zero(Fp& a) { a = fp_zero(); }
BN_HD Fp2 f_add(const Fp2& a, const Fp2& b) { return fp2_add(a, b); }
BN_HD Fp2 f_sub(const Fp2& a, const Fp2& b) { return fp2_sub(a, b); }
BN_HD Fp2 f_dbl(const Fp2& a) { return fp2_dbl(a); }
BN_HD Fp2 f_neg(const Fp2& a) { return fp2_neg(a); }
BN_HD Fp2 f_mul(const Fp2& a, const Fp2& b) { Fp2 z; fp2_mul(z, a, b); return z; }
BN_HD Fp2 f_sqr(const Fp2& a) { Fp2 z; fp2_sqr(z, a); return z; }
BN_HD Fp2 f_inv(const Fp2& a) { Fp2 z; fp2_inv(z, a); return z; }
BN_HD bool f_is_zero(const Fp2& a) { return fp2_is_zero(a); }
BN_HD void f_set_one(Fp2& a) { a = fp2_one(); }
BN_HD void f_set_zero(Fp2& a) { a = fp2_zero(); }

template <typename J>
BN_HD bool jac_is_inf(const J& p) { return f_is_zero(p.z); }

// a = 0 doubling (dbl-2009-l)
template <typename J>
BN_HD void jac_dbl(J& r, const J& p) {
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
BN_HD void jac_add_aff(J& r, const J& p, const A& q) {
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
template <typename A>
BN_HD bool aff_is_inf(const A& p) { return f_is_zero(p.x) && f_is_zero(p.y); }

// [s]base, s = 256-bit little-endian unsigned (8 x u32); signed 4-bit fixed windows over an
// on-the-fly table {1..8}*base kept in Jacobian-free affine form is overkill for a first version:
// plain left-to-right double-and-add with a 2-bit window is used until the GLV kernel lands.
template <typename J, typename A>
BN_HD void scalar_mul(A& out, const A& base, const uint32_t* s) {
  if (aff_is_inf(base)) { out = base; return; }
  // table: 1P (affine), 2P, 3P as Jacobian -> kept Jacobian and added via full addition is costly;
  // use affine 1P only (binary method).  Cost ~ 256 dbl + ~128 mixed add.
  J acc; f_set_zero(acc.x); f_set_zero(acc.y); f_set_zero(acc.z);
  int top = 255;
  while (top >= 0 && !((s[top >> 5] >> (top & 31)) & 1u)) top--;
  for (int i = top; i >= 0; i--) {
    jac_dbl(acc, acc);
    if ((s[i >> 5] >> (i & 31)) & 1u) jac_add_aff(acc, acc, base);
  }
  jac_to_aff(out, acc);
}
// affine + affine with gnark Add semantics (infinity operands, doubling, P + (-P))
template <typename J, typename A>
BN_HD void aff_add(A& out, const A& a, const A& b) {
  if (aff_is_inf(a)) { out = b; return; }
  if (aff_is_inf(b)) { out = a; return; }
  J t; t.x = a.x; t.y = a.y; f_set_one(t.z);
  jac_add_aff(t, t, b);
  jac_to_aff(out, t);
}

// GT.Exp: generic Fp12 square-and-multiply (no subgroup assumption), k = 256-bit LE; k == 0 -> 1
BN_HD void gt_exp(Fp12& out, const Fp12& x, const uint32_t* k) {
  Fp12 acc; fp12_set_one(acc);
  int top = 255;
  while (top >= 0 && !((k[top >> 5] >> (top & 31)) & 1u)) top--;
  if (top >= 0) {
    acc = x;
    for (int i = top - 1; i >= 0; i--) {
      fp12_sqr(acc, acc);
      if ((k[i >> 5] >> (i & 31)) & 1u) fp12_mul(acc, acc, x);
    }
  }
  out = acc;
}

}  // namespace bn254
