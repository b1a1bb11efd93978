// Host emulation of the DEVICE algorithms (csrc/*.cuh compiled by a plain C++ compiler; the PTX carry
// wrappers fall back to a C model).  TEST HARNESS ONLY: lets `-m "not gpu"` tests check the tower,
// Miller loop, final exponentiation and group code bit-for-bit against the oracle without a GPU.
// Never linked into libbn254_b200.so.
#include <cstring>
#include <cstddef>
#include "../../gopairingbasedcryptography_b200/csrc/curve.cuh"
#include "../../gopairingbasedcryptography_b200/csrc/hash_to_curve.cuh"
using namespace bn254;

template <typename T> static T ld(const void* p, size_t i) { T t; memcpy(&t, (const char*)p + i * sizeof(T), sizeof(T)); return t; }
template <typename T> static void st(void* p, size_t i, const T& t) { memcpy((char*)p + i * sizeof(T), &t, sizeof(T)); }

extern "C" {
void emu_fp_mul(const void* a, const void* b, size_t n, void* z) { for (size_t i = 0; i < n; i++) st(z, i, fp_mul(ld<Fp>(a, i), ld<Fp>(b, i))); }
void emu_fp_add(const void* a, const void* b, size_t n, void* z) { for (size_t i = 0; i < n; i++) st(z, i, fp_add(ld<Fp>(a, i), ld<Fp>(b, i))); }
void emu_fp_sub(const void* a, const void* b, size_t n, void* z) { for (size_t i = 0; i < n; i++) st(z, i, fp_sub(ld<Fp>(a, i), ld<Fp>(b, i))); }
void emu_fp_half(const void* a, size_t n, void* z) { for (size_t i = 0; i < n; i++) st(z, i, fp_half(ld<Fp>(a, i))); }
void emu_fp_inv(const void* a, size_t n, void* z) { for (size_t i = 0; i < n; i++) st(z, i, fp_inv(ld<Fp>(a, i))); }
void emu_multi_pair(const void* P, const void* Q, size_t n, size_t k, int mode, void* out) {
  for (size_t i = 0; i < n; i++) {
    Fp12 f; bool have = false;
    for (size_t base = 0; base < k; base += 4) {
      G1Aff p[4]; G2Aff q[4]; G2Proj T[4];
      int c = (int)((k - base) < 4 ? (k - base) : 4);
      for (int j = 0; j < c; j++) { p[j] = ld<G1Aff>(P, i * k + base + j); q[j] = ld<G2Aff>(Q, i * k + base + j); }
      if (!have) { miller_loop(f, p, q, T, c); have = true; } else { Fp12 g; miller_loop(g, p, q, T, c); fp12_mul(f, f, g); }
    }
    if (mode >= 1) final_exp(f, f);
    if (mode == 2) ((unsigned char*)out)[i] = fp12_is_one(f) ? 1 : 0; else st(out, i, f);
  }
}
void emu_final_exp(const void* in, size_t n, void* out) { for (size_t i = 0; i < n; i++) { Fp12 f = ld<Fp12>(in, i); final_exp(f, f); st(out, i, f); } }
void emu_g1_mul(const void* base, size_t stride, const void* s, size_t n, void* out) {
  for (size_t i = 0; i < n; i++) { G1Aff b = ld<G1Aff>(base, i * stride), r; uint32_t k[8]; memcpy(k, (const char*)s + 32 * i, 32); Fp beta = GLV_BETA; scalar_mul_glv<G1Jac, G1Aff>(r, b, k, beta); st(out, i, r); } }
void emu_g2_mul(const void* base, size_t stride, const void* s, size_t n, void* out) {
  for (size_t i = 0; i < n; i++) { G2Aff b = ld<G2Aff>(base, i * stride), r; uint32_t k[8]; memcpy(k, (const char*)s + 32 * i, 32); Fp beta = GLV_BETA_G2; scalar_mul_glv<G2Jac, G2Aff>(r, b, k, beta); st(out, i, r); } }
void emu_g1_add(const void* a, const void* b, size_t n, void* out) { for (size_t i = 0; i < n; i++) { G1Aff r; aff_add<G1Jac, G1Aff>(r, ld<G1Aff>(a, i), ld<G1Aff>(b, i)); st(out, i, r); } }
void emu_g2_add(const void* a, const void* b, size_t n, void* out) { for (size_t i = 0; i < n; i++) { G2Aff r; aff_add<G2Jac, G2Aff>(r, ld<G2Aff>(a, i), ld<G2Aff>(b, i)); st(out, i, r); } }
void emu_gt_exp(const void* x, size_t stride, const void* s, size_t n, void* out) {
  for (size_t i = 0; i < n; i++) { Fp12 b = ld<Fp12>(x, i * stride), r; uint32_t k[8]; memcpy(k, (const char*)s + 32 * i, 32); Fp12 tab[4]; gt_exp(r, b, k, tab); st(out, i, r); } }
void emu_gt_mul(const void* a, const void* b, size_t n, int div, void* out) {
  for (size_t i = 0; i < n; i++) { Fp12 x = ld<Fp12>(a, i), y = ld<Fp12>(b, i); if (div) fp12_inv(y, y); fp12_mul(x, x, y); st(out, i, x); } }
void emu_gt_sqr(const void* a, size_t n, int cyclo, void* out) {
  for (size_t i = 0; i < n; i++) { Fp12 x = ld<Fp12>(a, i); if (cyclo) fp12_cyclo_sqr(x, x); else fp12_sqr(x, x); st(out, i, x); } }
void emu_gt_frob(const void* a, size_t n, int k, void* out) { for (size_t i = 0; i < n; i++) { Fp12 x = ld<Fp12>(a, i); fp12_frob(x, x, k); st(out, i, x); } }
}

// ---- tower-VM programs executed on the host with lock-step round semantics (all lanes of a group load
// their operands before any lane stores), exactly what the warp does between two __syncwarp().
#include "../../gopairingbasedcryptography_b200/csrc/vm.cuh"
namespace {
const uint64_t kPair3[] = {
#include "../../gopairingbasedcryptography_b200/csrc/vm_prog_pair_k3.inc"
};
const uint64_t kMiller3[] = {
#include "../../gopairingbasedcryptography_b200/csrc/vm_prog_miller_k3.inc"
};
const uint64_t kFinalExp3[] = {
#include "../../gopairingbasedcryptography_b200/csrc/vm_prog_finalexp_k3.inc"
};
void run_rounds(Fp2* slots, const uint64_t* prog, int rounds, int K) {
  vm::SlotFile f; f.hot = reinterpret_cast<uint4*>(slots); f.cold = nullptr; f.nslots = 256; f.hot_stride = 0; f.pid = 0; f.cold_stride = 0; f.gpid = 0;
  for (int r = 0; r < rounds; r++) {
    int dst[8]; Fp2 val[8]; bool st[8];
    for (int j = 0; j < K; j++) st[j] = vm::exec_op(f, prog[(size_t)r * K + j], dst[j], val[j]);
    for (int j = 0; j < K; j++) if (st[j]) slots[dst[j]] = val[j];
  }
}
}
extern "C" {
// mode 0: pair (K=3), 1: miller only, 2: final exp only
void emu_vm(const void* in0, const void* in1, size_t n, int mode, void* out) {
  using namespace vm;
  for (size_t i = 0; i < n; i++) {
    Fp2 slots[256];
    const int* IN; const int* OUT; const uint64_t* prog; int rounds, K = 3, nin = 3;
    if (mode == 0) { IN = PAIR_K3_IN; OUT = PAIR_K3_OUT; prog = kPair3; rounds = PAIR_K3_ROUNDS; }
    else if (mode == 1) { IN = MILLER_K3_IN; OUT = MILLER_K3_OUT; prog = kMiller3; rounds = MILLER_K3_ROUNDS; }
    else { IN = FINALEXP_K3_IN; OUT = FINALEXP_K3_OUT; prog = kFinalExp3; rounds = FINALEXP_K3_ROUNDS; nin = 6; }
    if (nin == 3) {
      memcpy(&slots[IN[0]], (const char*)in0 + i * 64, 64);
      memcpy(&slots[IN[1]], (const char*)in1 + i * 128, 64);
      memcpy(&slots[IN[2]], (const char*)in1 + i * 128 + 64, 64);
    } else {
      for (int k = 0; k < 6; k++) memcpy(&slots[IN[k]], (const char*)in0 + i * 384 + k * 64, 64);
    }
    run_rounds(slots, prog, rounds, K);
    for (int k = 0; k < 6; k++) memcpy((char*)out + i * 384 + k * 64, &slots[OUT[k]], 64);
  }
}
}
extern "C" void emu_fp2_mul_lazy(const void* a, const void* b, size_t n, void* z) {
  for (size_t i = 0; i < n; i++) st(z, i, fp2_mul_lazy(ld<Fp2>(a, i), ld<Fp2>(b, i))); }
extern "C" void emu_fp2_mul(const void* a, const void* b, size_t n, void* z) {
  for (size_t i = 0; i < n; i++) st(z, i, fp2_mul_inl(ld<Fp2>(a, i), ld<Fp2>(b, i))); }
extern "C" void emu_fp_mul_wide_redc(const void* a, const void* b, size_t n, void* z) {
  for (size_t i = 0; i < n; i++) { uint32_t t[16]; fp_mul_wide(t, ld<Fp>(a, i), ld<Fp>(b, i)); st(z, i, fp_redc(t)); } }
extern "C" {
void emu_g1_mul_glv(const void* base, size_t stride, const void* s, size_t n, void* out) {
  for (size_t i = 0; i < n; i++) { G1Aff b = ld<G1Aff>(base, i * stride), r; uint32_t k[8]; memcpy(k, (const char*)s + 32 * i, 32);
    Fp beta = GLV_BETA; scalar_mul_glv<G1Jac, G1Aff>(r, b, k, beta); st(out, i, r); } }
void emu_g2_mul_glv(const void* base, size_t stride, const void* s, size_t n, void* out) {
  for (size_t i = 0; i < n; i++) { G2Aff b = ld<G2Aff>(base, i * stride), r; uint32_t k[8]; memcpy(k, (const char*)s + 32 * i, 32);
    Fp beta = GLV_BETA_G2; scalar_mul_glv<G2Jac, G2Aff>(r, b, k, beta); st(out, i, r); } }
void emu_g2_mul_gls4(const void* base, size_t stride, const void* s, size_t n, void* out) {
  for (size_t i = 0; i < n; i++) { G2Aff b = ld<G2Aff>(base, i * stride), r; uint32_t k[8]; memcpy(k, (const char*)s + 32 * i, 32);
    Fp2 slice[kGlsSliceFp2]; scalar_mul_gls4(r, b, k, slice); st(out, i, r); } }
// fixed base: table built with the GLV routine exactly as the engine does (scalars d << 8w)
void emu_g1_mul_fixed(const void* base1, const void* s, size_t n, void* out) {
  static G1Aff table[kFixedWindows * kFixedEntries];
  G1Aff b = ld<G1Aff>(base1, 0); Fp beta = GLV_BETA;
  for (int w = 0; w < kFixedWindows; w++) for (int d = 1; d <= kFixedEntries; d++) {
    uint32_t k[8] = {0, 0, 0, 0, 0, 0, 0, 0}; k[w >> 2] = (uint32_t)d << ((w & 3) * 8);
    scalar_mul_glv<G1Jac, G1Aff>(table[w * kFixedEntries + d - 1], b, k, beta); }
  for (size_t i = 0; i < n; i++) { uint32_t k[8]; memcpy(k, (const char*)s + 32 * i, 32); G1Aff r; scalar_mul_fixed<G1Jac, G1Aff>(r, table, k); st(out, i, r); } }
}
extern "C" void emu_gt_cyclo_exp(const void* x, size_t stride, const void* s, size_t n, void* out) {
  for (size_t i = 0; i < n; i++) { Fp12 b = ld<Fp12>(x, i * stride), r; uint32_t k[8]; memcpy(k, (const char*)s + 32 * i, 32); Fp12 tab[16]; gt_cyclo_exp(r, b, k, tab); st(out, i, r); } }
// line tables: precompute for the m points of Q, then out[i] = FE(prod_j lines_j evaluated at P[i*m+j])
extern "C" void emu_multi_pair_lines(const void* P, const void* Q, size_t n, size_t m, void* out) {
  Fp2* table = new Fp2[m * kLinesPerPoint * 3];
  unsigned char* qs = new unsigned char[m];
  for (size_t j = 0; j < m; j++) { G2Aff q = ld<G2Aff>(Q, j); qs[j] = g2_is_inf(q); if (!qs[j]) g2_precompute_lines(q, table + j * kLinesPerPoint * 3); }
  for (size_t i = 0; i < n; i++) {
    Fp12 f; fp12_set_one(f);
    int s = 0;
    for (int it = ATE_NAF_LEN - 2; it >= -2; it--) {
      if (it >= 0 && it != ATE_NAF_LEN - 2) fp12_sqr(f, f);
      int reps = (it >= 0 && ATE_NAF[it]) ? 2 : 1;
      for (int r = 0; r < reps; r++, s++)
        for (size_t j0 = 0; j0 < m; j0 += 8) {  // chunks of 8 table points, lines two at a time: the kernel's own step function
          int cnt = (int)(m - j0 < 8 ? m - j0 : 8);
          G1Aff p[8];
          unsigned skip = 0;
          for (int j = 0; j < cnt; j++) { p[j] = ld<G1Aff>(P, i * m + j0 + j); if (g1_is_inf(p[j]) || qs[j0 + j]) skip |= 1u << j; }
          Fp2 sc[kScratchSlots];
          miller_lines_step(f, p, table, (int)j0, cnt, skip, s, sc);
        }
    }
    final_exp(f, f);
    st(out, i, f);
  }
  delete[] table; delete[] qs;
}

// hash-to-curve device code on the host (same sources as k_hash_to_curve)
extern "C" void emu_hash_to_g1(const unsigned char* msg, size_t len, const unsigned char* dst, size_t dst_len, void* out) {
  G1Aff r; hash_to_g1(r, msg, len, dst, (uint32_t)dst_len); st(out, 0, r); }
extern "C" void emu_hash_to_g2(const unsigned char* msg, size_t len, const unsigned char* dst, size_t dst_len, void* out) {
  G2Aff r; hash_to_g2(r, msg, len, dst, (uint32_t)dst_len); st(out, 0, r); }
extern "C" void emu_fp_is_square(const void* a, size_t n, int* out) { for (size_t i = 0; i < n; i++) out[i] = fp_is_square(ld<Fp>(a, i)) ? 1 : 0; }
extern "C" void emu_hash_to_field(const unsigned char* msg, size_t len, const unsigned char* dst, size_t dst_len, void* out) {
  Fp u[4]; hash_to_field<4>(u, msg, len, dst, (uint32_t)dst_len); for (int i = 0; i < 4; i++) st(out, i, u[i]); }

extern "C" void emu_fp2_mul_xi(const void* a, size_t n, void* z) { for (size_t i = 0; i < n; i++) st(z, i, fp2_mul_xi_i(ld<Fp2>(a, i))); }

// ---- warp-VM programs (wvmgen.py / wvm.cuh) on the host: 32 lanes per round, all lanes load before any lane stores ----
#include "../../gopairingbasedcryptography_b200/csrc/wvm.cuh"
#include "../../gopairingbasedcryptography_b200/csrc/wvm_prog_meta.cuh"
namespace {
const uint32_t kWvmMillerH[] = {
#include "../../gopairingbasedcryptography_b200/csrc/wvm_prog_miller.inc"
};
const uint32_t kWvmMiller2H[] = {
#include "../../gopairingbasedcryptography_b200/csrc/wvm_prog_miller2.inc"
};
const uint32_t kWvmFinalExpH[] = {
#include "../../gopairingbasedcryptography_b200/csrc/wvm_prog_finalexp.inc"
};
void wvm_rounds(Fp* slots, const uint32_t* prog, int rounds) {
  for (int r = 0; r < rounds; r++) {
    unsigned dst[32]; Fp val[32]; bool st_[32];
    for (int j = 0; j < 32; j++) {
      uint4 w0, w1;
      memcpy(&w0, prog + ((size_t)r * 32 + j) * 8, 16);
      memcpy(&w1, prog + ((size_t)r * 32 + j) * 8 + 4, 16);
      st_[j] = wvm::exec_op(slots, w0, w1, 15, dst[j], val[j]);
    }
    for (int j = 0; j < 32; j++) if (st_[j]) slots[dst[j]] = val[j];
  }
}
}
extern "C" {
// mode 0: Miller loop, 1: pairing (Miller program, then the final-exponentiation program), 2: final exponentiation
void emu_wvm(const void* in0, const void* in1, size_t n, int mode, void* out) {
  using namespace wvm;
  for (size_t i = 0; i < n; i++) {
    Fp slots[1024];
    memset(slots, 0, sizeof(slots));
    Fp v[12];
    if (mode != 2) {
      for (int k = 0; k < MILLER_NCONST; k++) slots[MILLER_CONST_SLOT[k]] = MILLER_CONST_VAL[k];
      memcpy(&slots[MILLER_IN[0]], (const char*)in0 + i * 64, 32);
      memcpy(&slots[MILLER_IN[1]], (const char*)in0 + i * 64 + 32, 32);
      for (int k = 0; k < 4; k++) memcpy(&slots[MILLER_IN[2 + k]], (const char*)in1 + i * 128 + k * 32, 32);
      wvm_rounds(slots, kWvmMillerH, MILLER_ROUNDS);
      for (int k = 0; k < 12; k++) v[k] = slots[MILLER_OUT[k]];
    } else {
      for (int k = 0; k < 12; k++) memcpy(&v[k], (const char*)in0 + i * 384 + k * 32, 32);
    }
    if (mode >= 1) {
      for (int k = 0; k < FINALEXP_NCONST; k++) slots[FINALEXP_CONST_SLOT[k]] = FINALEXP_CONST_VAL[k];
      for (int k = 0; k < 12; k++) slots[FINALEXP_IN[k]] = v[k];
      wvm_rounds(slots, kWvmFinalExpH, FINALEXP_ROUNDS);
      for (int k = 0; k < 12; k++) v[k] = slots[FINALEXP_OUT[k]];
    }
    memcpy((char*)out + i * 384, v, 384);
  }
}
// the two-pair Miller program (k_wvm_miller2): in0 = P[2n], in1 = Q[2n] -> n Miller products
void emu_wvm_miller2(const void* in0, const void* in1, size_t n, void* out) {
  using namespace wvm;
  for (size_t i = 0; i < n; i++) {
    Fp slots[1024];
    memset(slots, 0, sizeof(slots));
    for (int k = 0; k < MILLER2_NCONST; k++) slots[MILLER2_CONST_SLOT[k]] = MILLER2_CONST_VAL[k];
    for (int j = 0; j < 2; j++) {
      memcpy(&slots[MILLER2_IN[6 * j]], (const char*)in0 + i * 128 + j * 64, 32);
      memcpy(&slots[MILLER2_IN[6 * j + 1]], (const char*)in0 + i * 128 + j * 64 + 32, 32);
      for (int k = 0; k < 4; k++) memcpy(&slots[MILLER2_IN[6 * j + 2 + k]], (const char*)in1 + i * 256 + j * 128 + k * 32, 32);
    }
    wvm_rounds(slots, kWvmMiller2H, MILLER2_ROUNDS);
    for (int k = 0; k < 12; k++) memcpy((char*)out + i * 384 + k * 32, &slots[MILLER2_OUT[k]], 32);
  }
}
// the LIN reduction on its own: any 9-limb total below 256 p
void emu_wvm_lin_reduce(const void* v9, size_t n, void* out) {
  for (size_t i = 0; i < n; i++) { uint32_t v[9]; memcpy(v, (const char*)v9 + i * 36, 36); st(out, i, wvm::lin_reduce9(v)); }
}
// one LIN op: out = sum_i c[i] * s[i] mod p over nterms Montgomery operands (coefficients as int32)
void emu_wvm_lin(const void* s, const int* c, size_t nterms, void* out) {
  wvm::LinAcc acc; wvm::lin_init(acc);
  for (size_t i = 0; i < nterms; i++) wvm::lin_acc(acc, ld<Fp>(s, i), c[i]);
  st(out, 0, wvm::lin_finish(acc));
}
void emu_wvm_inv(const void* a, size_t n, void* z) { for (size_t i = 0; i < n; i++) st(z, i, wvm::fp_inv_wvm(ld<Fp>(a, i))); }
}
