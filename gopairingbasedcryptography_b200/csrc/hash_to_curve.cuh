// Hash-to-curve for BN254 G1 / G2 on the device: SHA-256 expand_message_xmd, hash_to_field (L = 48), the
// Shallue-van de Woestijne map (RFC 9380 section 6.6.1 / F.1), point addition and, for G2, cofactor clearing with
// the psi endomorphism  [x0]Q + psi([3x0]Q) + psi^2([x0]Q) + psi^3(Q).
// Replaces (reference side): gnark-crypto v0.19.0 ecc/bn254/hash_to_g1.go {HashToG1, MapToCurve1},
// hash_to_g2.go {HashToG2, MapToCurve2}, g2.go ClearCofactor, fp/hash.go Hash + ExpandMsgXmd, reached from
// hash/hash_to.go:113-119,169-175,203-209,271-277 (ToG1 / BytesToG1 / ToG2 / BytesToG2), i.e.
// signature/bls01_signature/bls_signature.go:60,73, ibe/bf01_ibe/bf01_ibe.go:130,158, dabe/lw11_dabe.go:96,177.
// One message per thread; control flow is lane-uniform except for the message length.
#pragma once
#include "curve.cuh"

namespace bn254 {

// Field products use the by-value helpers of curve.cuh (f_mul / f_sqr / f_inv).
BN_HD Fp h_mul(const Fp& a, const Fp& b) { return f_mul(a, b); }
BN_HD Fp h_sqr(const Fp& a) { return f_sqr(a); }
BN_HD Fp h_inv(const Fp& a) { return f_inv(a); }
BN_HD Fp2 h_mul(const Fp2& a, const Fp2& b) { return f_mul(a, b); }
BN_HD Fp2 h_sqr(const Fp2& a) { return f_sqr(a); }
BN_HD Fp2 h_inv(const Fp2& a) { return f_inv(a); }

// ---- SHA-256 (FIPS 180-4), byte-streaming ------------------------------------------------------------------
BN_CONST uint32_t SHA256_K[64] = {
    0x428a2f98u, 0x71374491u, 0xb5c0fbcfu, 0xe9b5dba5u, 0x3956c25bu, 0x59f111f1u, 0x923f82a4u, 0xab1c5ed5u,
    0xd807aa98u, 0x12835b01u, 0x243185beu, 0x550c7dc3u, 0x72be5d74u, 0x80deb1feu, 0x9bdc06a7u, 0xc19bf174u,
    0xe49b69c1u, 0xefbe4786u, 0x0fc19dc6u, 0x240ca1ccu, 0x2de92c6fu, 0x4a7484aau, 0x5cb0a9dcu, 0x76f988dau,
    0x983e5152u, 0xa831c66du, 0xb00327c8u, 0xbf597fc7u, 0xc6e00bf3u, 0xd5a79147u, 0x06ca6351u, 0x14292967u,
    0x27b70a85u, 0x2e1b2138u, 0x4d2c6dfcu, 0x53380d13u, 0x650a7354u, 0x766a0abbu, 0x81c2c92eu, 0x92722c85u,
    0xa2bfe8a1u, 0xa81a664bu, 0xc24b8b70u, 0xc76c51a3u, 0xd192e819u, 0xd6990624u, 0xf40e3585u, 0x106aa070u,
    0x19a4c116u, 0x1e376c08u, 0x2748774cu, 0x34b0bcb5u, 0x391c0cb3u, 0x4ed8aa4au, 0x5b9cca4fu, 0x682e6ff3u,
    0x748f82eeu, 0x78a5636fu, 0x84c87814u, 0x8cc70208u, 0x90befffau, 0xa4506cebu, 0xbef9a3f7u, 0xc67178f2u};

struct Sha256 {
  uint32_t h[8];
  uint32_t w[16];  // current block, big-endian words
  uint32_t fill;   // bytes in the block
  uint64_t total;  // bytes absorbed
};
BN_HD uint32_t sha_rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
BN_HD void sha256_init(Sha256& s) {
  s.h[0] = 0x6a09e667u; s.h[1] = 0xbb67ae85u; s.h[2] = 0x3c6ef372u; s.h[3] = 0xa54ff53au;
  s.h[4] = 0x510e527fu; s.h[5] = 0x9b05688cu; s.h[6] = 0x1f83d9abu; s.h[7] = 0x5be0cd19u;
  for (int i = 0; i < 16; i++) s.w[i] = 0;
  s.fill = 0; s.total = 0;
}
BN_NOINLINE void sha256_compress(Sha256& s) {
  uint32_t w[64];
  for (int i = 0; i < 16; i++) w[i] = s.w[i];
  for (int i = 16; i < 64; i++) {
    uint32_t s0 = sha_rotr(w[i - 15], 7) ^ sha_rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
    uint32_t s1 = sha_rotr(w[i - 2], 17) ^ sha_rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
    w[i] = w[i - 16] + s0 + w[i - 7] + s1;
  }
  uint32_t a = s.h[0], b = s.h[1], c = s.h[2], d = s.h[3], e = s.h[4], f = s.h[5], g = s.h[6], hh = s.h[7];
  for (int i = 0; i < 64; i++) {
    uint32_t S1 = sha_rotr(e, 6) ^ sha_rotr(e, 11) ^ sha_rotr(e, 25);
    uint32_t ch = (e & f) ^ (~e & g);
    uint32_t t1 = hh + S1 + ch + SHA256_K[i] + w[i];
    uint32_t S0 = sha_rotr(a, 2) ^ sha_rotr(a, 13) ^ sha_rotr(a, 22);
    uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
    uint32_t t2 = S0 + mj;
    hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
  }
  s.h[0] += a; s.h[1] += b; s.h[2] += c; s.h[3] += d; s.h[4] += e; s.h[5] += f; s.h[6] += g; s.h[7] += hh;
  for (int i = 0; i < 16; i++) s.w[i] = 0;
  s.fill = 0;
}
BN_HD void sha256_byte(Sha256& s, uint32_t b) {
  s.w[s.fill >> 2] |= (b & 0xffu) << (24 - 8 * (s.fill & 3));
  s.fill++; s.total++;
  if (s.fill == 64) sha256_compress(s);
}
BN_HD void sha256_bytes(Sha256& s, const uint8_t* p, size_t n) { for (size_t i = 0; i < n; i++) sha256_byte(s, p[i]); }
BN_HD void sha256_final(Sha256& s, uint32_t* digest /* 8 big-endian words */) {
  uint64_t bits = s.total * 8;
  sha256_byte(s, 0x80);
  while (s.fill != 56) sha256_byte(s, 0);
  s.w[14] = (uint32_t)(bits >> 32); s.w[15] = (uint32_t)bits;
  sha256_compress(s);
  for (int i = 0; i < 8; i++) digest[i] = s.h[i];
}
BN_HD void sha256_words(Sha256& s, const uint32_t* w, int n) { for (int i = 0; i < n; i++) for (int k = 0; k < 4; k++) sha256_byte(s, w[i] >> (24 - 8 * k)); }

// expand_message_xmd(msg, dst, 48 * COUNT) -> COUNT field elements in Montgomery form (gnark fp.Hash: each 48-byte
// chunk is a big-endian integer reduced mod p).  COUNT = 2 (G1) or 4 (G2): ell = 3 or 6 digest blocks.
template <int COUNT>
BN_HD void hash_to_field(Fp* u, const uint8_t* msg, size_t msg_len, const uint8_t* dst, uint32_t dst_len) {
  constexpr int LEN = 48 * COUNT, ELL = LEN / 32;
  uint32_t b0[8], bi[8], out[ELL * 8];
  Sha256 s;
  sha256_init(s);
  for (int i = 0; i < 64; i++) sha256_byte(s, 0);  // Z_pad
  sha256_bytes(s, msg, msg_len);
  sha256_byte(s, LEN >> 8); sha256_byte(s, LEN & 0xff); sha256_byte(s, 0);
  sha256_bytes(s, dst, dst_len); sha256_byte(s, dst_len);
  sha256_final(s, b0);
  for (int i = 1; i <= ELL; i++) {
    sha256_init(s);
    if (i == 1) sha256_words(s, b0, 8);
    else { uint32_t x[8]; for (int k = 0; k < 8; k++) x[k] = b0[k] ^ bi[k]; sha256_words(s, x, 8); }
    sha256_byte(s, i);
    sha256_bytes(s, dst, dst_len); sha256_byte(s, dst_len);
    sha256_final(s, bi);
    for (int k = 0; k < 8; k++) out[(i - 1) * 8 + k] = bi[k];
  }
  // out = LEN bytes as big-endian words; element j = words [12j, 12j+12): most significant word first
  for (int j = 0; j < COUNT; j++) {
    Fp acc = fp_zero();
    for (int c = 0; c < 3; c++) {  // c = 0: least significant 128 bits = the last four words of the chunk
      Fp piece = fp_zero();
      for (int k = 0; k < 4; k++) piece.l[k] = out[12 * j + 11 - 4 * c - k];
      const Fp& K = c == 0 ? H2F_K0 : (c == 1 ? H2F_K1 : H2F_K2);
      acc = fp_add(acc, h_mul(piece, K));
    }
    u[j] = acc;
  }
}

// ---- field helpers: fixed-exponent powers, squareness, square roots, sgn0 -----------------------------------
BN_HD Fp fp_pow_fixed(const Fp& base, const uint32_t* e) { return fp_pow_win(base, e); }  // base^e, e = 8 constant limbs (curve.cuh: sliding window)
// Quadratic character without a single field multiplication: the binary Jacobi symbol (a / p) -- strip the trailing
// zeros of a (flip the sign for an odd count when p = +-3 mod 8), make a the larger of the two odd numbers (flip when
// both are 3 mod 4: reciprocity), subtract, repeat until a = 0; the symbol is the sign when the other number ended at 1.
// About 190 rounds of ~50 shift / compare / subtract instructions on 8 limbs instead of the ~313 Montgomery products
// (~59 000 instructions) of the Euler ladder a^((p-1)/2).  The argument may stay in Montgomery form: R = 2^256 is a
// square, so (a R / p) = (a / p).  Data-dependent trip count: the lanes of a warp differ by a few rounds only.
#ifndef BN254_EULER_LADDER
BN_HD int fp_ctz32(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __ffs((int)x) - 1;
#else
  return __builtin_ctz(x);
#endif
}
BN_NOINLINE int fp_jacobi(Fp x) {
  uint32_t a[8], n[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { a[i] = x.l[i]; n[i] = p_limb(i); }
  uint32_t t = 0;  // parity of the sign flips
  for (;;) {
    uint32_t any = a[0] | a[1] | a[2] | a[3] | a[4] | a[5] | a[6] | a[7];
    if (!any) break;
    while (a[0] == 0) {  // whole zero limbs: 32 halvings, no flip
#pragma unroll
      for (int i = 0; i < 7; i++) a[i] = a[i + 1];
      a[7] = 0;
    }
    const int z = fp_ctz32(a[0]);
    if (z) {
#pragma unroll
      for (int i = 0; i < 7; i++) a[i] = (a[i] >> z) | (a[i + 1] << (32 - z));
      a[7] >>= z;
      const uint32_t n8 = n[0] & 7u;
      t ^= (uint32_t)(z & 1) & ((n8 == 3u || n8 == 5u) ? 1u : 0u);
    }
    // a is odd now; order the pair so that a >= n
    bool lt = false;
#pragma unroll
    for (int i = 7; i >= 0; i--) {  // lexicographic compare from the top limb
      bool ne = a[i] != n[i];
      lt = ne ? (a[i] < n[i]) : lt;
      if (ne) break;
    }
    if (lt) {
#pragma unroll
      for (int i = 0; i < 8; i++) { uint32_t w = a[i]; a[i] = n[i]; n[i] = w; }
      t ^= ((a[0] & 3u) == 3u && (n[0] & 3u) == 3u) ? 1u : 0u;
    }
    uint64_t bw = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { uint64_t d = (uint64_t)a[i] - n[i] - bw; a[i] = (uint32_t)d; bw = (d >> 32) & 1u; }
  }
  uint32_t rest = (n[0] ^ 1u) | n[1] | n[2] | n[3] | n[4] | n[5] | n[6] | n[7];
  return rest ? 0 : (t ? -1 : 1);
}
BN_HD bool fp_is_square(const Fp& a) { return fp_is_zero(a) || fp_jacobi(a) == 1; }
#else
BN_HD bool fp_is_square(const Fp& a) {  // Euler: a^((p-1)/2) is 1 (or a = 0)
  Fp t = fp_pow_fixed(a, FP_PM1H);
  return fp_is_zero(a) || fp_eq(t, fp_one());
}
#endif
BN_HD Fp fp_sqrt(const Fp& a) { return fp_pow_fixed(a, FP_PP1Q); }  // valid when a is a square
BN_HD uint32_t fp_sgn0(const Fp& a) { Fp raw = FP_RAW_ONE; return h_mul(a, raw).l[0] & 1u; }  // parity of the regular form
BN_HD Fp fp2_norm(const Fp2& a) { return fp_add(h_sqr(a.a0), h_sqr(a.a1)); }
BN_HD bool fp2_is_square(const Fp2& a) { return fp_is_square(fp2_norm(a)); }
BN_HD uint32_t fp2_sgn0(const Fp2& a) {
  uint32_t s0 = fp_sgn0(a.a0), z0 = fp_is_zero(a.a0) ? 1u : 0u;
  return s0 | (z0 & fp_sgn0(a.a1));
}
BN_HD Fp fp_sel(bool c, const Fp& a, const Fp& b) {  // c ? a : b, as a mask blend
  uint32_t m = 0u - (c ? 1u : 0u);
  Fp z;
#pragma unroll
  for (int i = 0; i < 8; i++) z.l[i] = (a.l[i] & m) | (b.l[i] & ~m);
  return z;
}
// some square root of a square in Fp2 (complex method); which of the two is irrelevant, the map fixes the sign
BN_NOINLINE Fp2 fp2_sqrt(Fp2 a) {
  Fp2 r;
  if (fp_is_zero(a.a1)) {
    bool sq = fp_is_square(a.a0);
    Fp t = fp_sqrt(sq ? a.a0 : fp_neg(a.a0));
    r.a0 = sq ? t : fp_zero();
    r.a1 = sq ? fp_zero() : t;
    return r;
  }
  // Complex method with ONE more ladder instead of three (an Euler test, a square root and an inversion).  With
  // n = sqrt(norm a) the candidates c+- = (a0 +- n) / 2 satisfy c+ c- = -a1^2 / 4, so exactly one of them is a square
  // (p = 3 mod 4: -1 is a non-residue).  t = c+^((p-3)/4) and x = t c+ = c+^((p+1)/4) give t x = c+^((p-1)/2) = +-1:
  //   c+ square:      x^2 = c+,   1/x = t    ->  root (x, a1 / (2 x)) = (x, a1 t / 2)           (the previous code's value)
  //   c+ non-residue: x^2 = -c+,  1/x = -t,  c- = a1^2 / (4 x^2)  ->  root (a1 / (2 x), x) = (-a1 t / 2, x)
  // In the second case this may be the negative of the root the three-ladder code returned; the only caller fixes the
  // sign by sgn0 right after (map_to_curve_g2), so the mapped point is bit-identical.
  Fp n = fp_sqrt(fp2_norm(a));
  Fp half = FP_HALF;
  Fp c = h_mul(fp_add(a.a0, n), half);
  Fp t = fp_pow_fixed(c, FP_PM3Q);
  Fp x = h_mul(t, c);
  bool sq = fp_eq(h_sqr(x), c);
  Fp a1t = h_mul(a.a1, h_mul(t, half));
  r.a0 = fp_sel(sq, x, fp_neg(a1t));
  r.a1 = fp_sel(sq, a1t, x);
  return r;
}
BN_HD Fp2 fp2_sel(bool c, const Fp2& a, const Fp2& b) { Fp2 z; z.a0 = fp_sel(c, a.a0, b.a0); z.a1 = fp_sel(c, a.a1, b.a1); return z; }

// ---- SVDW maps (steps as in RFC 9380 F.1 / gnark MapToCurve1, MapToCurve2) ------------------------------------
// The map in two halves around its one inversion, so that the two maps of a message can SHARE it (Montgomery's trick:
// 1 / a and 1 / b from one inversion of a b and three products -- one fixed-exponent ladder less per message on the
// one-thread-per-message kernels).  inv0 semantics kept: a zero operand gets the inverse 0 and does not disturb the other.
BN_HD Fp g1_curve_rhs(const Fp& x) { Fp b = CURVE_B; return fp_add(h_mul(h_sqr(x), x), b); }
BN_HD void svdw_pre_g1(Fp& tv1, Fp& tv2, Fp& prod, const Fp& u) {
  Fp c1 = H2C_G1_C1, one = fp_one();
  Fp t = h_mul(h_sqr(u), c1);
  tv2 = fp_add(one, t);
  tv1 = fp_sub(one, t);
  prod = h_mul(tv1, tv2);
}
BN_NOINLINE void svdw_post_g1(G1Aff& out, const Fp& u, const Fp& tv1, const Fp& tv2, const Fp& tv3) {
  Fp c2 = H2C_G1_C2, c3 = H2C_G1_C3, c4 = H2C_G1_C4, Z = H2C_G1_Z;
  Fp tv4 = h_mul(h_mul(h_mul(u, tv1), tv3), c3);
  Fp x1 = fp_sub(c2, tv4);
  bool e1 = fp_is_square(g1_curve_rhs(x1));
  Fp x2 = fp_add(c2, tv4);
  bool e2 = fp_is_square(g1_curve_rhs(x2)) && !e1;
  Fp x3 = h_mul(h_sqr(tv2), tv3);
  x3 = fp_add(h_mul(h_sqr(x3), c4), Z);
  Fp x = fp_sel(e1, x1, x3);
  x = fp_sel(e2, x2, x);
  Fp y = fp_sqrt(g1_curve_rhs(x));
  if (fp_sgn0(u) != fp_sgn0(y)) y = fp_neg(y);
  out.x = x; out.y = y;
}
BN_NOINLINE void map_to_curve_g1(G1Aff& out, const Fp& u) {
  Fp tv1, tv2, prod;
  svdw_pre_g1(tv1, tv2, prod, u);
  svdw_post_g1(out, u, tv1, tv2, h_inv(prod));  // inv0
}
// both maps of one message, one inversion
template <typename INV = InvThread>
BN_HD void map_to_curve_g1_x2(G1Aff& q0, G1Aff& q1, const Fp& u0, const Fp& u1) {
  Fp a1, a2, pa, b1, b2, pb, one = fp_one();
  svdw_pre_g1(a1, a2, pa, u0);
  svdw_pre_g1(b1, b2, pb, u1);
  bool za = fp_is_zero(pa), zb = fp_is_zero(pb);
  Fp sa = fp_sel(za, one, pa), sb = fp_sel(zb, one, pb);
  Fp inv = INV::inv(h_mul(sa, sb));
  Fp ia = fp_sel(za, fp_zero(), h_mul(inv, sb)), ib = fp_sel(zb, fp_zero(), h_mul(inv, sa));
  svdw_post_g1(q0, u0, a1, a2, ia);
  svdw_post_g1(q1, u1, b1, b2, ib);
}
BN_HD Fp2 g2_curve_rhs(const Fp2& x) { Fp2 b = TWIST_B; return fp2_add_i(h_mul(h_sqr(x), x), b); }
BN_HD void svdw_pre_g2(Fp2& tv1, Fp2& tv2, Fp2& prod, const Fp2& u) {
  Fp2 c1 = H2C_G2_C1, one = fp2_one();
  Fp2 t = h_mul(h_sqr(u), c1);
  tv2 = fp2_add_i(one, t);
  tv1 = fp2_sub_i(one, t);
  prod = h_mul(tv1, tv2);
}
BN_NOINLINE void svdw_post_g2(G2Aff& out, const Fp2& u, const Fp2& tv1, const Fp2& tv2, const Fp2& tv3) {
  Fp2 c2 = H2C_G2_C2, c3 = H2C_G2_C3, c4 = H2C_G2_C4, Z = H2C_G2_Z;
  Fp2 tv4 = h_mul(h_mul(h_mul(u, tv1), tv3), c3);
  Fp2 x1 = fp2_sub_i(c2, tv4);
  bool e1 = fp2_is_square(g2_curve_rhs(x1));
  Fp2 x2 = fp2_add_i(c2, tv4);
  bool e2 = fp2_is_square(g2_curve_rhs(x2)) && !e1;
  Fp2 x3 = h_mul(h_sqr(tv2), tv3);
  x3 = fp2_add_i(h_mul(h_sqr(x3), c4), Z);
  Fp2 x = fp2_sel(e1, x1, x3);
  x = fp2_sel(e2, x2, x);
  Fp2 y = fp2_sqrt(g2_curve_rhs(x));
  if (fp2_sgn0(u) != fp2_sgn0(y)) y = fp2_neg_i(y);
  out.x = x; out.y = y;
}
BN_NOINLINE void map_to_curve_g2(G2Aff& out, const Fp2& u) {
  Fp2 tv1, tv2, prod;
  svdw_pre_g2(tv1, tv2, prod, u);
  svdw_post_g2(out, u, tv1, tv2, h_inv(prod));  // inv0 (fp_inv(0) = 0)
}
template <typename INV = InvThread>
BN_HD void map_to_curve_g2_x2(G2Aff& q0, G2Aff& q1, const Fp2& u0, const Fp2& u1) {
  Fp2 a1, a2, pa, b1, b2, pb, one = fp2_one();
  svdw_pre_g2(a1, a2, pa, u0);
  svdw_pre_g2(b1, b2, pb, u1);
  bool za = fp2_is_zero(pa), zb = fp2_is_zero(pb);
  Fp2 sa = fp2_sel(za, one, pa), sb = fp2_sel(zb, one, pb);
  Fp2 inv = INV::inv(h_mul(sa, sb));
  Fp2 ia = fp2_sel(za, fp2_zero(), h_mul(inv, sb)), ib = fp2_sel(zb, fp2_zero(), h_mul(inv, sa));
  svdw_post_g2(q0, u0, a1, a2, ia);
  svdw_post_g2(q1, u1, b1, b2, ib);
}

// psi = twist o Frobenius o untwist on Jacobian coordinates: (X, Y, Z) -> (conj X g12, conj Y g13, conj Z)
BN_HD void g2_psi(G2Jac& r, const G2Jac& p) {
  G2Jac t;
  Fp2 g2c = GAMMA1[2], g3c = GAMMA1[3];
  t.x = h_mul(fp2_conj_i(p.x), g2c);
  t.y = h_mul(fp2_conj_i(p.y), g3c);
  t.z = fp2_conj_i(p.z);
  r = t;
}
// gnark G2Jac.ClearCofactor (Fuentes-Castaneda et al., section 6.1)
BN_NOINLINE void g2_clear_cofactor(G2Jac& r, const G2Jac& q) {
  // [x0]Q over the width-3 signed digits of x0 (the constant the GT exponentiation by x0 uses: 18 non-zero digits in
  // {+-1, +-3}) with 3Q precomputed: 63 doublings + 20 additions instead of 63 + 28 over the plain bits -- the same group
  // element, so the same affine point after normalisation; the digits are constants, the loop is lane-uniform.
  G2Jac q3;
  jac_dbl(q3, q); jac_add(q3, q3, q);
  G2Jac xq; f_set_zero(xq.x); f_set_zero(xq.y); f_set_zero(xq.z);
  for (int i = X0_NAF3_LEN - 1; i >= 0; i--) {
    jac_dbl(xq, xq);
    const int dgt = X0_NAF3[i];
    if (dgt) {
      G2Jac m = (dgt == 1 || dgt == -1) ? q : q3;
      if (dgt < 0) m.y = fp2_neg_i(m.y);
      jac_add(xq, xq, m);
    }
  }
  G2Jac p1, p2, p3, acc;
  jac_dbl(p1, xq); jac_add(p1, p1, xq); g2_psi(p1, p1);
  g2_psi(p2, xq); g2_psi(p2, p2);
  g2_psi(p3, q); g2_psi(p3, p3); g2_psi(p3, p3);
  jac_add(acc, xq, p1); jac_add(acc, acc, p2); jac_add(acc, acc, p3);
  r = acc;
}

// INV = InvCta (kernels with every thread of the CTA alive): the shared SVDW inversion and the final normalisation go
// through ONE Fermat ladder per CTA instead of one per thread (curve.cuh); every thread calls each exactly once.
template <typename INV = InvThread>
BN_HD void hash_to_g1(G1Aff& out, const uint8_t* msg, size_t len, const uint8_t* dst, uint32_t dst_len) {
  Fp u[2];
  hash_to_field<2>(u, msg, len, dst, dst_len);
  G1Aff q0, q1;
  map_to_curve_g1_x2<INV>(q0, q1, u[0], u[1]);
  // gnark Add semantics on two finite points: doubling when equal, infinity (0, 0) when opposite (inv(0) = 0)
  G1Jac t; t.x = q0.x; t.y = q0.y; t.z = fp_one();
  jac_add_aff(t, t, q1);
  jac_to_aff_inv<INV>(out, t);
}
template <typename INV = InvThread>
BN_HD void hash_to_g2(G2Aff& out, const uint8_t* msg, size_t len, const uint8_t* dst, uint32_t dst_len) {
  Fp u[4];
  hash_to_field<4>(u, msg, len, dst, dst_len);
  G2Aff q0, q1;
  Fp2 u0, u1;
  u0.a0 = u[0]; u0.a1 = u[1]; u1.a0 = u[2]; u1.a1 = u[3];
  map_to_curve_g2_x2<INV>(q0, q1, u0, u1);
  G2Jac s; s.x = q0.x; s.y = q0.y; s.z = fp2_one();
  jac_add_aff(s, s, q1);
  g2_clear_cofactor(s, s);
  jac_to_aff_inv<INV>(out, s);
}

}  // namespace bn254
