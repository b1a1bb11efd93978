// debug: per-step outputs of the G1 SVDW map on the device
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
#define BN254_OOL_ADDS
#define BN254_OOL_FPMUL
#include "../../gopairingbasedcryptography_b200/csrc/hash_to_curve.cuh"
using namespace bn254;
__device__ __noinline__ void map_dbg(G1Aff& out, const Fp& u, Fp* dbg, int* fl) {
  Fp c1 = H2C_G1_C1, c2 = H2C_G1_C2, c3 = H2C_G1_C3, c4 = H2C_G1_C4, Z = H2C_G1_Z, one = fp_one();
  Fp tv1 = f_mul(f_sqr(u), c1);
  Fp tv2 = fp_add(one, tv1);
  tv1 = fp_sub(one, tv1);
  Fp tv3 = f_inv(f_mul(tv1, tv2));  // inv0
  Fp tv4 = f_mul(f_mul(f_mul(u, tv1), tv3), c3);
  Fp x1 = fp_sub(c2, tv4);
  bool e1 = fp_is_square(g1_curve_rhs(x1));
  Fp x2 = fp_add(c2, tv4);
  Fp gx2 = g1_curve_rhs(x2);
  Fp tt; fp_pow_fixed(tt, gx2, FP_PM1H);
  dbg[8] = gx2; dbg[9] = tt;
  bool sq2 = fp_is_square(gx2);
  fl[3] = sq2; fl[4] = fp_eq(tt, fp_one()); fl[5] = fp_is_zero(gx2);
  bool e2 = sq2 && !e1;
  Fp x3 = f_mul(f_sqr(tv2), tv3);
  x3 = fp_add(f_mul(f_sqr(x3), c4), Z);
  fl[6] = e1; fl[7] = e2;
  dbg[0] = x1; dbg[1] = x2; dbg[2] = x3;
  Fp x = fp_sel(e1, x1, x3);
  dbg[3] = x;
  x = fp_sel(e2, x2, x);
  dbg[4] = x;
  Fp y = fp_sqrt(g1_curve_rhs(x));
  if (fp_sgn0(u) != fp_sgn0(y)) y = fp_neg(y);
  out.x = x; out.y = y;
}
__global__ void k(const unsigned char* msg, int len, const unsigned char* dst, int dlen, Fp* out, int* flags) {
  Fp u[2];
  hash_to_field<2>(u, msg, len, dst, dlen);
  out[0] = u[0]; out[1] = u[1];
  for (int j = 0; j < 2; j++) {
    Fp c1 = H2C_G1_C1, c2 = H2C_G1_C2, c3 = H2C_G1_C3, one = fp_one();
    Fp tv1 = f_mul(f_sqr(u[j]), c1);
    Fp tv2 = fp_add(one, tv1);
    tv1 = fp_sub(one, tv1);
    Fp tv3 = f_inv(f_mul(tv1, tv2));
    Fp tv4 = f_mul(f_mul(f_mul(u[j], tv1), tv3), c3);
    Fp x1 = fp_sub(c2, tv4);
    Fp gx1 = g1_curve_rhs(x1);
    Fp t; fp_pow_fixed(t, gx1, FP_PM1H);
    out[2 + 6 * j] = tv3; out[3 + 6 * j] = x1; out[4 + 6 * j] = gx1; out[5 + 6 * j] = t;
    flags[4 * j] = fp_is_square(gx1);
    flags[4 * j + 1] = fp_sgn0(u[j]);
    Fp y = fp_sqrt(gx1);
    out[6 + 6 * j] = y;
    if (j == 1) {
      Fp c4 = H2C_G1_C4, Z = H2C_G1_Z;
      Fp x2 = fp_add(c2, tv4);
      Fp gx2 = g1_curve_rhs(x2);
      flags[3] = fp_is_square(gx2);
      Fp x3 = f_mul(f_sqr(tv2), tv3);
      Fp x3b = fp_add(f_mul(f_sqr(x3), c4), Z);
      out[2] = x2; out[3] = gx2; out[4] = x3; out[5] = x3b; out[6] = c4;
    }
    G1Aff r; map_to_curve_g1(r, u[j]);
    if (j == 1) { G1Aff r2; map_dbg(r2, u[j], out + 8, flags + 5); out[13] = r2.x; }
    if (j == 0) out[7] = r.x; else out[15] = r.x;
    flags[4 * j + 2] = fp_sgn0(r.y);
  }
}
int main() {
  const char* msg = "abc"; const char* dst = "Hash Bytes To Element In G1";
  unsigned char *dm, *dd; Fp* dout; int* dfl;
  cudaMalloc(&dm, 16); cudaMalloc(&dd, 64); cudaMalloc(&dout, 32 * 32); cudaMalloc(&dfl, 128); cudaMemset(dfl, 0xff, 128);
  cudaMemcpy(dm, msg, 3, cudaMemcpyHostToDevice); cudaMemcpy(dd, dst, strlen(dst), cudaMemcpyHostToDevice);
  k<<<1, 1>>>(dm, 3, dd, (int)strlen(dst), dout, dfl);
  Fp h[32]; int fl[32];
  cudaError_t e = cudaMemcpy(h, dout, sizeof(h), cudaMemcpyDeviceToHost); cudaMemcpy(fl, dfl, sizeof(fl), cudaMemcpyDeviceToHost);
  printf("err %s\n", cudaGetErrorString(e));
  for (int i = 0; i < 20; i++) { printf("%d ", i); for (int j = 7; j >= 0; j--) printf("%08x", h[i].l[j]); printf("\n"); }
  for (int i = 0; i < 16; i++) printf("flag %d = %d\n", i, fl[i]);
}
