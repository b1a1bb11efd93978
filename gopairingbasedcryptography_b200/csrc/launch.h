// Host-visible launchers of the kernel translation units (k_pairing.cu, k_group.cu, k_gt.cu, k_hash.cu, k_vm.cu).
// engine.cu (the host runtime behind include/bn254_b200.h) sees only this header: plain pointers, sizes and a
// stream.  Every launcher enqueues on `s` and returns; errors surface through cudaGetLastError() in the caller.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include <atomic>

namespace bn254 { namespace launch {

// kernel launches issued by this library in this process (bn254_launch_count): every <<<>>> below counts one
extern std::atomic<uint64_t> g_launches;
#define BN_LAUNCH (::bn254::launch::g_launches.fetch_add(1, std::memory_order_relaxed))

constexpr int kBlockThreads = 128;        // CTA size of every thread kernel
constexpr int kMpChunk = 8;               // pairs per thread in split multi-pairings
#ifndef BN254_LINES_CHUNK
#define BN254_LINES_CHUNK 8
#endif
constexpr int kLinesChunk = BN254_LINES_CHUNK;  // table points per thread in the line-table Miller kernel
constexpr int kLinesPerPoint = 65 + 21 + 2;  // lines of one G2 point's Miller schedule
constexpr size_t kLineBytes = 3 * 64;     // (r0, r1, r2) Fp2 coefficients
constexpr int kFixedWindows = 32, kFixedEntries = 255;  // fixed-base tables: 8-bit windows, digits 1..255
constexpr size_t kGtExpTable = 4, kGtCycloTable = 16;   // per-thread Fp12 table entries of the GT ladders
constexpr int kMsmWindowBits = 8;         // shared-point MSM: per-point tables of [d * 2^(8w)] P
constexpr int kMsmWindows = 32;

// per-family one-time setup (dynamic shared-memory attributes); call once per process/device after cudaSetDevice
cudaError_t pairing_init();
cudaError_t group_init();
cudaError_t gt_init();
cudaError_t hash_init();

// ---- pairing family ---------------------------------------------------------------------------------------------
void pair(const void* P, const void* Q, size_t n, void* out, cudaStream_t s);
// mode 0: Miller product, 1: + final exponentiation, 2: check (one byte per product); 1 <= k <= 2 * kMpChunk
void multi_pair(int mode, const void* P, const void* Q, size_t n, int k, void* out, cudaStream_t s);
void mp_partial(const void* P, const void* Q, size_t n, int k, int nchunks, void* partial, cudaStream_t s);
void mp_combine(int mode, const void* partial, size_t n, int nchunks, void* out, cudaStream_t s);
void g2_lines(const void* Q, size_t m, void* table, uint8_t* qskip, cudaStream_t s);
void miller_lines(const void* P, const void* table, const uint8_t* qskip, size_t n, int m, int nchunks, void* partial, cudaStream_t s);
void final_exp(const void* in, size_t n, void* out, cudaStream_t s);
void check2_fixed_g1(const void* P01, const void* Q0, const void* Q1, size_t n, uint8_t* ok, cudaStream_t s);
void pack_check2(const void* P01, const void* Q0, const void* Q1, size_t n, void* P, void* Q, cudaStream_t s);
void gt_is_one(const void* x, size_t n, uint8_t* ok, cudaStream_t s);
int pairing_wave_threads(int sms);  // threads of one full wave of the pairing kernels on `sms` SMs

// ---- group family (g = 1: G1, 2: G2) ----------------------------------------------------------------------------
void scalar_mul(int g, const void* base, size_t base_stride, const void* scalars, size_t n, void* out, cudaStream_t s);
// G2 by the 4-dimensional GLS ladder; scratch = g2_gls_scratch_bytes(n) bytes of device memory, private to the launch
size_t g2_gls_scratch_bytes(size_t n);
void scalar_mul_g2_gls(const void* base, size_t base_stride, const void* scalars, size_t n, void* out, void* scratch, cudaStream_t s);
void fixed_mul(int g, const void* table, const void* scalars, size_t n, void* out, cudaStream_t s);
void aff_add(int g, const void* a, const void* b, size_t n, void* out, cudaStream_t s);
void subset_sum(int g, const void* U, int m, const uint8_t* sel, size_t n, void* out, cudaStream_t s);
// byte-window form: builds table (ceil(m/8) x 256 affine points, caller scratch) then ceil(m/8) additions per selector
void subset_sum_tab(int g, const void* U, int m, const uint8_t* sel, size_t n, void* out, void* table, cudaStream_t s);
void segment_sum(int g, const void* pts, size_t groups, int len, int chunk, void* out, cudaStream_t s);
size_t subset_sum_max_bytes();
// shared-point MSM: out[v] = sum_j [s[v*len + j]] P_j over per-point window tables (built by msm_tables)
// zs, pf: scratch of len * 32 * 255 field elements each (32 B G1 / 64 B G2), free after the launch completes
void msm_tables(int g, const void* pts, size_t len, void* tables, void* zs, void* pf, cudaStream_t s);
void msm_partial(int g, const void* tables, const void* scalars, size_t nvec, size_t len, int chunk, void* partial, cudaStream_t s);
void msm_partial_tree(int g, const void* tables, const void* scalars, size_t nvec, size_t len, int chunk, void* partial, cudaStream_t s);  // one partial per CTA
void jac_tree(int g, const void* in, size_t groups, int len, void* out_a, cudaStream_t s);  // len a power of two <= kBlockThreads
// Jacobian partial sums (96 B G1 / 192 B G2 each): out_j (Jacobian) or, on the last pass, out_a (canonical affine)
void jac_sum(int g, const void* in, size_t groups, int len, int chunk, void* out_j, void* out_a, cudaStream_t s);
void neg_points(int g, const void* in, size_t n, void* out, cudaStream_t s);

// ---- GT family --------------------------------------------------------------------------------------------------
void gt_exp(int cyclo, const void* x, size_t x_stride, const void* k, size_t n, void* out, void* tabmem, cudaStream_t s);
void gt_fixed_exp(const void* table, const void* k, size_t n, void* out, cudaStream_t s);
void gt_mul(int mode, const void* a, size_t a_stride, const void* b, size_t b_stride, size_t n, void* out, cudaStream_t s);  // mode 0: a*b, 1: a/b, 2: a*conj(b)
void fp_mul(const void* a, const void* b, size_t n, void* out, cudaStream_t s);
int gt_wave_threads(int sms);

// ---- hash-to-curve ----------------------------------------------------------------------------------------------
void hash_to_curve(int g, const uint8_t* msgs, const uint64_t* off, size_t n, const uint8_t* dst, uint32_t dst_len, void* out, cudaStream_t s);

// ---- Fr feeders (k_fr.cu): fr.Element = 32 B Montgomery, gnark layout -------------------------------------------
cudaError_t fr_poly_from_roots(const void* roots, size_t n, void* coeffs /* n + 1 */, cudaStream_t s);
void fr_quotient_coeffs(const void* f /* n + 1 */, size_t n, const void* ids, size_t nvec, void* out_scalars /* nvec x n, regular form */, cudaStream_t s);
void fr_to_scalars(const void* in, size_t n, void* out, cudaStream_t s);
void fr_lagrange_basis_host(const void* s, size_t n, const void* x, void* out);  // CPU, one inversion for the set
void fr_to_scalars_host(const void* in, size_t n, void* out);

// ---- lane-group (tower VM) kernels ------------------------------------------------------------------------------
enum { kVmPair = 0, kVmMiller = 1, kVmFinalExp = 2, kVmMiller2 = 3 };  // Miller2: product of TWO Miller loops per item (warp-VM only)
constexpr int kVmPairingsPerCta = 40;  // 4 warps x 10 lane groups of 3 (k_vm.cu asserts it)
cudaError_t vm_prepare(int* blocks_per_sm /* [3] */);
size_t vm_cold_bytes(int sms, const int* blocks_per_sm);
void vm_run(int prog, const void* a, const void* b, size_t n, void* out, void* cold, int sms, const int* blocks_per_sm, cudaStream_t s);

// ---- warp-VM kernels (k_wvm.cu): one warp per item, the latency path; same program ids as the lane-group kernels ----
cudaError_t wvm_prepare(int* blocks_per_sm /* [4] */);
int wvm_items_per_cta();
void wvm_run(int prog, const void* a, const void* b, size_t n, void* out, int sms, const int* blocks_per_sm, cudaStream_t s);

} }  // namespace bn254::launch
