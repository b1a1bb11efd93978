/* bn254_b200.h -- C ABI of the B200-native batched BN254 pairing engine (libbn254_b200.so).
 *
 * This is the drop-in boundary for the hot path of mmsyan/GoPairingBasedCryptography: the calls the
 * schemes make into github.com/consensys/gnark-crypto v0.19.0 `ecc/bn254` (go.mod:5), re-exposed
 * as BATCH entry points.  A Go package binds these through cgo (INTEGRATION.md shows the stubs);
 * in this repository the Python mirror gopairingbasedcryptography_b200/bn254.py binds them through
 * ctypes.  There is no CPU fallback: every entry point fails with BN254_ERR_CUDA when no sm_100
 * device is usable.
 *
 * Data layout = gnark's in-memory layout, so Go values can be passed with unsafe.Pointer, no
 * conversion:
 *   fp.Element / fr.Element  32 B   4 x u64 little-endian limbs, Montgomery form (R = 2^256), < p
 *   G1Affine {X,Y fp}        64 B   point at infinity = all zero
 *   G2Affine {X,Y E2{A0,A1}} 128 B  point at infinity = all zero
 *   GT = E12{C0,C1 E6{B0,B1,B2 E2}}  384 B
 *   scalar                   32 B   little-endian unsigned integer, REGULAR form (big.Int value);
 *                                   the Go shim reduces fr.Element -> big.Int exactly as the
 *                                   reference does with x.BigInt(new(big.Int)).
 * Arrays are contiguous AoS.  "_dev" variants take device pointers already resident in HBM and a
 * cudaStream_t (as void*), enqueue asynchronously and do not synchronize; the others take host
 * pointers and return when the results are in `out`: buffers in page-locked memory (bn254_host_alloc,
 * cudaHostRegister) are copied to / from the device directly, chunk by chunk on the context's two
 * streams; pageable buffers go through the context's pinned staging area.
 *
 * Return value: 0 on success, negative BN254_ERR_* otherwise.  Thread-safety: a context is
 * internally locked -- every call holds the context's mutex from its first table lookup to its last launch (and,
 * for host-buffer calls, until the results are in `out`); use one context per GPU (or several per GPU for
 * concurrency).  Table handles (bn254_lines, bn254_fixed_base, bn254_msm_table) are immutable after create().
 */
#ifndef BN254_B200_H
#define BN254_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define BN254_OK 0
#define BN254_ERR_INVALID_SIZES (-1) /* gnark: errors.New("invalid inputs sizes") from Pair/MillerLoop/PairingCheck */
#define BN254_ERR_CUDA (-2)          /* no device, launch or copy failure; see bn254_last_error */
#define BN254_ERR_OOM (-3)
#define BN254_ERR_BAD_ARG (-4)

#define BN254_G1_BYTES 64
#define BN254_G2_BYTES 128
#define BN254_GT_BYTES 384
#define BN254_SCALAR_BYTES 32

typedef struct bn254_ctx bn254_ctx;

/* lifecycle: one context per GPU ordinal.  bn254_ctx_destroy also releases the device memory of every table / line
 * handle created on the context that is still alive; destroying such a handle AFTERWARDS is allowed (a garbage-collected
 * host gives no destruction order) and only frees the handle itself.  Using it in any other call is an error. */
int bn254_ctx_create(int device, bn254_ctx** out);
void bn254_ctx_destroy(bn254_ctx* ctx);
const char* bn254_last_error(bn254_ctx* ctx); /* text of the last failure on this context */
int bn254_device_count(void);
/* pinned host memory for zero-staging transfers (optional) */
void* bn254_host_alloc(size_t bytes);
void bn254_host_free(void* p);
/* number of kernel launches issued through this context so far (bench.py's gpu_launches) */
uint64_t bn254_launch_count(bn254_ctx* ctx);
int bn254_sm_count(bn254_ctx* ctx);

/* Device memory and streams for hosts without a CUDA binding of their own (the Go package: per-GPU async streams,
 * intermediates kept in HBM between *_dev calls).  Streams are cudaStream_t passed as void*; NULL is the default
 * stream.  bn254_dev_upload enqueues (page-locked sources copy asynchronously); bn254_dev_download returns when the
 * bytes are in h_dst. */
int bn254_dev_alloc(bn254_ctx*, size_t bytes, void** d_out);
int bn254_dev_free(bn254_ctx*, void* d);
int bn254_dev_upload(bn254_ctx*, void* d_dst, const void* h_src, size_t bytes, void* stream);
int bn254_dev_download(bn254_ctx*, void* h_dst, const void* d_src, size_t bytes, void* stream);
int bn254_stream_create(bn254_ctx*, void** stream_out);
int bn254_stream_destroy(bn254_ctx*, void* stream);
int bn254_stream_sync(bn254_ctx*, void* stream);

/* bn254.Generators()  [31 call sites, e.g. signature/bls01_signature/bls_signature.go:32] */
void bn254_generators(void* g1_aff_64, void* g2_aff_128);

/* n independent bn254.Pair([]G1Affine{P[i]}, []G2Affine{Q[i]})
 * [access/tree/access_tree_node.go:106,110; cpabe/bsw07/bsw07_cpabe.go:184;
 *  ibe/waters05_ibe/waters05_ibe.go:214,259,262; bibe/afp25_bibe/afp25_bibe.go:395-403] */
int bn254_pair_batch(bn254_ctx*, const void* P, const void* Q, size_t n, void* out_gt);
int bn254_pair_batch_dev(bn254_ctx*, const void* dP, const void* dQ, size_t n, void* d_out_gt, void* stream);

/* n products bn254.Pair(P[i*k..i*k+k), Q[i*k..i*k+k)): one Miller product, ONE final exponentiation
 * each; pairs containing infinity are skipped; k == 0 -> BN254_ERR_INVALID_SIZES.
 * [the fused multi-pairing shape of access/tree/access_tree_node.go:96-164 + bsw07_cpabe.go:172-195] */
int bn254_multi_pair_batch(bn254_ctx*, const void* P, const void* Q, size_t n, size_t k, void* out_gt);
int bn254_multi_pair_batch_dev(bn254_ctx*, const void* dP, const void* dQ, size_t n, size_t k, void* d_out_gt, void* stream);

/* Precomputed G2 line tables for FIXED G2 points (a user key: bsw07 usk.dj / usk.djPrime / usk.d,
 * cpabe/bsw07/bsw07_cpabe.go:97-131; public parameters of the IBE schemes).  gnark's analogue is
 * PrecomputeLines / MillerLoopFixedQ.  create() runs the G2 side of the Miller schedule once per point
 * (88 lines x 192 B each, kept on the GPU); bn254_multi_pair_lines_batch then computes, for n rows of m G1 points,
 * out[i] = Pair(P[i*m .. i*m+m), Q[0..m)) with no G2 arithmetic and one final exponentiation per row --
 * bit-identical to bn254_multi_pair_batch on the same operands. */
typedef struct bn254_lines bn254_lines;
int bn254_g2_lines_create(bn254_ctx*, const void* Q, size_t m, bn254_lines** out);
void bn254_g2_lines_destroy(bn254_lines*);
size_t bn254_g2_lines_count(const bn254_lines*);
int bn254_multi_pair_lines_batch(bn254_ctx*, const void* P, const bn254_lines* lines, size_t n, void* out_gt);
int bn254_multi_pair_lines_batch_dev(bn254_ctx*, const void* dP, const bn254_lines* lines, size_t n, void* d_out_gt, void* stream);

/* n x bn254.PairingCheck(P[i*k..], Q[i*k..]) -> ok[i] in {0,1}
 * [signature/bls01_signature/bls_signature.go:81-84] */
int bn254_pairing_check_batch(bn254_ctx*, const void* P, const void* Q, size_t n, size_t k, uint8_t* ok);
int bn254_pairing_check_batch_dev(bn254_ctx*, const void* dP, const void* dQ, size_t n, size_t k, uint8_t* d_ok, void* stream);

/* bn254.MillerLoop / bn254.FinalExponentiation (0 call sites in the reference; named by north_star).
 * The raw Miller value is defined only up to factors the final exponentiation kills;
 * final_exp(miller_loop(P,Q)) == pair(P,Q) bit-exactly. */
int bn254_miller_loop_batch(bn254_ctx*, const void* P, const void* Q, size_t n, size_t k, void* out_gt);
int bn254_final_exp_batch(bn254_ctx*, const void* in_gt, size_t n, void* out_gt);
int bn254_miller_loop_batch_dev(bn254_ctx*, const void* dP, const void* dQ, size_t n, size_t k, void* d_out_gt, void* stream);
int bn254_final_exp_batch_dev(bn254_ctx*, const void* d_in_gt, size_t n, void* d_out_gt, void* stream);

/* (*G1Affine).ScalarMultiplication(&base[i], s[i]) / (*G2Affine).ScalarMultiplication
 * [bls_signature.go:63; waters05_ibe.go:237; bsw07_cpabe.go:149,160; afp25_bibe_utils.go:48,51] */
int bn254_g1_mul_batch(bn254_ctx*, const void* base, const void* scalars, size_t n, void* out);
int bn254_g2_mul_batch(bn254_ctx*, const void* base, const void* scalars, size_t n, void* out);
/* ScalarMultiplicationBase and any other fixed point: ONE base, n scalars
 * [bls_signature.go:45; waters05_ibe.go:224; bsw07_cpabe.go:69,157; afp25_bibe.go:160] */
int bn254_g1_mul_base_batch(bn254_ctx*, const void* base1, const void* scalars, size_t n, void* out);
int bn254_g2_mul_base_batch(bn254_ctx*, const void* base1, const void* scalars, size_t n, void* out);
int bn254_g1_mul_batch_dev(bn254_ctx*, const void* d_base, size_t base_stride_elems, const void* d_scalars, size_t n, void* d_out, void* stream);
int bn254_g2_mul_batch_dev(bn254_ctx*, const void* d_base, size_t base_stride_elems, const void* d_scalars, size_t n, void* d_out, void* stream);

/* Explicit fixed-base handles: an immutable 32 x 255 window table of ONE
 * base -- [d << 8w] base for G1 / G2, base^(d << 8w) for GT -- owned by the caller and usable from any thread for as
 * long as it lives.  The *_mul_base_batch / *_exp_base_batch entry points above keep a small per-context cache of
 * such tables instead (keyed by the base's bytes); use a handle when a base outlives many calls (public parameters:
 * waters05 g1, e(g1,g2)^alpha [ibe/waters05_ibe/waters05_ibe.go:219,224]; bsw07 g2, g2^alpha [bsw07_cpabe.go:69-83]). */
#define BN254_GROUP_G1 1
#define BN254_GROUP_G2 2
#define BN254_GROUP_GT 3
typedef struct bn254_fixed_base bn254_fixed_base;
int bn254_fixed_base_create(bn254_ctx*, int group, const void* base, bn254_fixed_base** out);
void bn254_fixed_base_destroy(bn254_fixed_base*);
int bn254_fixed_base_group(const bn254_fixed_base*);
int bn254_g1_fixed_mul_batch(bn254_ctx*, const bn254_fixed_base*, const void* scalars, size_t n, void* out);
int bn254_g2_fixed_mul_batch(bn254_ctx*, const bn254_fixed_base*, const void* scalars, size_t n, void* out);
int bn254_gt_fixed_exp_batch(bn254_ctx*, const bn254_fixed_base*, const void* k, size_t n, void* out);
int bn254_g1_fixed_mul_batch_dev(bn254_ctx*, const bn254_fixed_base*, const void* d_scalars, size_t n, void* d_out, void* stream);
int bn254_g2_fixed_mul_batch_dev(bn254_ctx*, const bn254_fixed_base*, const void* d_scalars, size_t n, void* d_out, void* stream);
int bn254_gt_fixed_exp_batch_dev(bn254_ctx*, const bn254_fixed_base*, const void* d_k, size_t n, void* d_out, void* stream);

/* Shared-point multi-scalar multiplication: out[v] = sum_j [s[v*len + j]] P_j for nvec coefficient vectors over the
 * SAME len points -- the shape of bibe/afp25_bibe/afp25_bibe_utils.go:45-55 (computeG1PolynomialTau over
 * g1, [tau]1 .. [tau^B]1, once per Digest and once per Decrypt) and bibe/gwww25_bibe/gwww25_bibe_utils.go:40-50 (G2).
 * create() builds per-point byte-window tables once per public parameter set (len x 32 x 255 affine points:
 * 0.5 MB per G1 point, 1 MB per G2 point); a term then costs 32 mixed additions and no doubling.
 * scalars: nvec x len x 32 B (any 256-bit value); out: nvec canonical affine points. */
typedef struct bn254_msm_table bn254_msm_table;
int bn254_msm_table_create(bn254_ctx*, int group, const void* points, size_t len, bn254_msm_table** out);
void bn254_msm_table_destroy(bn254_msm_table*);
size_t bn254_msm_table_len(const bn254_msm_table*);
int bn254_msm_batch(bn254_ctx*, const bn254_msm_table*, const void* scalars, size_t nvec, void* out);
int bn254_msm_batch_dev(bn254_ctx*, const bn254_msm_table*, const void* d_scalars, size_t nvec, void* d_out, void* stream);

/* (*G1Affine).Add / (*G2Affine).Add, canonical affine result, gnark semantics for infinity,
 * doubling and P + (-P)  [waters05_ibe.go:227-233; bsw07_cpabe.go:104,119] */
int bn254_g1_add_batch(bn254_ctx*, const void* a, const void* b, size_t n, void* out);
int bn254_g2_add_batch(bn254_ctx*, const void* a, const void* b, size_t n, void* out);
int bn254_g1_add_batch_dev(bn254_ctx*, const void* da, const void* db, size_t n, void* d_out, void* stream);
int bn254_g2_add_batch_dev(bn254_ctx*, const void* da, const void* db, size_t n, void* d_out, void* stream);
/* (*G1Affine).Neg / (*G2Affine).Neg on device-resident arrays (host callers negate 32 bytes themselves) */
int bn254_g1_neg_batch_dev(bn254_ctx*, const void* d_in, size_t n, void* d_out, void* stream);
int bn254_g2_neg_batch_dev(bn254_ctx*, const void* d_in, size_t n, void* d_out, void* stream);

/* Waters hash  out[i] = U[0] + sum_{j<m : bit j of sel_i} U[j+1]  for n selector strings of ceil(m/8) bytes, bit j =
 * bit (7 - j%8) of byte j/8 (the identity-vector order of ibe/waters05_ibe/waters05_ibe.go:302-313).  Replaces the
 * loop of affine Adds at waters05_ibe.go:227-233 (one inversion per Add) by one Jacobian sum per identity. */
int bn254_g1_subset_sum_batch(bn254_ctx*, const void* U_m_plus_1, size_t m, const void* sel, size_t n, void* out);
int bn254_g2_subset_sum_batch(bn254_ctx*, const void* U_m_plus_1, size_t m, const void* sel, size_t n, void* out);
int bn254_g1_subset_sum_batch_dev(bn254_ctx*, const void* dU, size_t m, const void* d_sel, size_t n, void* d_out, void* stream);
int bn254_g2_subset_sum_batch_dev(bn254_ctx*, const void* dU, size_t m, const void* d_sel, size_t n, void* d_out, void* stream);
/* out[g] = points[g*len] + ... + points[g*len+len-1]: the Add chain of bibe/afp25_bibe/afp25_bibe_utils.go:45-55
 * (after bn254_g1_mul_batch on the terms) and of gwww25's G2-side MSM. */
int bn254_g1_sum_batch(bn254_ctx*, const void* points, size_t groups, size_t len, void* out);
int bn254_g2_sum_batch(bn254_ctx*, const void* points, size_t groups, size_t len, void* out);
int bn254_g1_sum_batch_dev(bn254_ctx*, const void* d_points, size_t groups, size_t len, void* d_out, void* stream);
int bn254_g2_sum_batch_dev(bn254_ctx*, const void* d_points, size_t groups, size_t len, void* d_out, void* stream);

/* (*GT).Exp(x[i], k[i]) with k >= 0 (the Go shim inverts x for negative k as gnark does); generic
 * Fp12 exponentiation, no subgroup assumption; k == 0 -> 1
 * [access/tree/access_tree_node.go:156; waters05_ibe.go:219; bsw07_cpabe.go:80,146] */
int bn254_gt_exp_batch(bn254_ctx*, const void* x, const void* k, size_t n, void* out);
int bn254_gt_exp_base_batch(bn254_ctx*, const void* x1, const void* k, size_t n, void* out);
int bn254_gt_exp_batch_dev(bn254_ctx*, const void* d_x, size_t x_stride_elems, const void* d_k, size_t n, void* d_out, void* stream);
/* Same result as bn254_gt_exp_batch when x lies in GT proper, the order-r subgroup (any Pair output, or a product,
 * quotient or power of Pair outputs -- every GT.Exp base in the reference's non-test code):
 * 2-dimensional GLV split of the exponent (x^lambda = conj(frobenius^2 x)), Granger-Scott squarings, joint fixed
 * windows: ~2.5x less work.  Undefined for other Fp12 elements (including cyclotomic elements of order not
 * dividing r). */
int bn254_gt_cyclo_exp_batch(bn254_ctx*, const void* x, const void* k, size_t n, void* out);
int bn254_gt_cyclo_exp_base_batch(bn254_ctx*, const void* x1, const void* k, size_t n, void* out);
int bn254_gt_cyclo_exp_batch_dev(bn254_ctx*, const void* d_x, size_t x_stride_elems, const void* d_k, size_t n, void* d_out, void* stream);
/* (*GT).Mul / (*GT).Div  [access_tree_node.go:114,157; bsw07_cpabe.go:189-190] */
int bn254_gt_mul_batch(bn254_ctx*, const void* a, const void* b, size_t n, void* out);
int bn254_gt_div_batch(bn254_ctx*, const void* a, const void* b, size_t n, void* out);
/* strides in elements: 0 broadcasts one element over the batch, 1 walks an array */
int bn254_gt_mul_batch_dev(bn254_ctx*, const void* da, size_t a_stride, const void* db, size_t b_stride, size_t n, void* d_out, void* stream);
int bn254_gt_div_batch_dev(bn254_ctx*, const void* da, size_t a_stride, const void* db, size_t b_stride, size_t n, void* d_out, void* stream);
/* a / b for b in GT proper (a pairing output, or a product / quotient / power of pairing outputs -- the divisor at
 * every Div of the reference's decryption flows: bsw07_cpabe.go:189-190, waters05_ibe.go:269-274, afp25_bibe.go:407-413):
 * b is unitary there, so b^-1 = conj(b) and the quotient is ONE Fp12 product instead of an Fp12 inversion (~800 dependent
 * Fp products on one thread) plus a product.  Same bytes as bn254_gt_div_batch for such b; undefined for other b. */
int bn254_gt_cyclo_div_batch(bn254_ctx*, const void* a, const void* b, size_t n, void* out);
int bn254_gt_cyclo_div_batch_dev(bn254_ctx*, const void* da, size_t a_stride, const void* db, size_t b_stride, size_t n, void* d_out, void* stream);

/* ok[i] = PairingCheck({P0, P1}, {Q0[i], Q1[i]}) with the two G1 points shared by the batch: BLS verification
 * [signature/bls01_signature/bls_signature.go:71-89: P0 = pk, P1 = -g1, Q0[i] = H(m_i), Q1[i] = sigma_i].
 * P01: 128 B (P0 then P1); Q0, Q1: n x 128 B; ok: n bytes. */
int bn254_pairing_check2_fixed_g1_batch(bn254_ctx*, const void* P01, const void* Q0, const void* Q1, size_t n, uint8_t* ok);
int bn254_pairing_check2_fixed_g1_batch_dev(bn254_ctx*, const void* dP01, const void* dQ0, const void* dQ1, size_t n, uint8_t* d_ok, void* stream);

/* bn254.HashToG1(msg, dst) / bn254.HashToG2(msg, dst) for n messages at once
 * [hash/hash_to.go:113-119 ToG1, 169-175 BytesToG1, 203-209 ToG2, 271-277 BytesToG2; callers
 *  signature/bls01_signature/bls_signature.go:60,73, ibe/bf01_ibe/bf01_ibe.go:130,158, dabe/lw11_dabe.go:96,177,
 *  bibe/afp25_bibe/afp25_bibe_utils.go:10-12].  RFC 9380 hash_to_curve as gnark configures it for BN254:
 * expand_message_xmd(SHA-256), 48 bytes per field element, Shallue-van de Woestijne map (Z = 1 on G1, Z = u on G2),
 * sum of the two mapped points, G2 cofactor cleared with the psi endomorphism.  Everything runs on the GPU.
 * msgs: the n messages concatenated; offsets: n + 1 byte offsets into msgs (message i = [offsets[i], offsets[i+1]));
 * dst: domain separation tag, at most 255 bytes (longer -> BN254_ERR_BAD_ARG, where gnark returns an error);
 * out: n canonical affine points (64 B / 128 B each). */
int bn254_hash_to_g1_batch(bn254_ctx*, const uint8_t* msgs, const uint64_t* offsets, size_t n, const uint8_t* dst, size_t dst_len, void* out);
int bn254_hash_to_g2_batch(bn254_ctx*, const uint8_t* msgs, const uint64_t* offsets, size_t n, const uint8_t* dst, size_t dst_len, void* out);
/* device-resident messages / offsets / points; dst stays a host pointer (at most 255 bytes, copied at call time) */
int bn254_hash_to_g1_batch_dev(bn254_ctx*, const uint8_t* d_msgs, const uint64_t* d_offsets, size_t n, const uint8_t* dst, size_t dst_len, void* d_out, void* stream);
int bn254_hash_to_g2_batch_dev(bn254_ctx*, const uint8_t* d_msgs, const uint64_t* d_offsets, size_t n, const uint8_t* dst, size_t dst_len, void* d_out, void* stream);

/* Scalar-field feeders.  fr.Element = 32 B, 4 x u64 little-endian limbs, Montgomery form,
 * as gnark stores it; "scalars" = the big.Int value as 32 little-endian bytes (what every *_mul / *_exp entry takes).
 *  - bn254_fr_poly_from_roots: c_0..c_n of f(X) = prod (X - root_i)
 *    [bibe/afp25_bibe/afp25_bibe_utils.go:14-43 computePolynomialCoeffs; gwww25_bibe_utils.go], n <= 3000;
 *  - bn254_fr_quotient_coeffs: for every id, the n coefficients of f(X) / (X - id) by O(n) synthetic division
 *    (the reference re-expands the polynomial without the root: O(n^2) per Decrypt, afp25_bibe.go:369-383), written
 *    as scalars in bn254_msm_batch's layout out[v * n + k];
 *  - bn254_fr_lagrange_basis (host only): Delta_{s_i,S}(x) for all i with ONE inversion
 *    [utils/compute_lagrange_basis.go:8-30: one inversion per factor]; fr.Element in, fr.Element out;
 *  - bn254_fr_to_scalars: x.BigInt(new(big.Int)) for an array. */
int bn254_fr_poly_from_roots(bn254_ctx*, const void* roots, size_t n, void* coeffs_n_plus_1);
int bn254_fr_poly_from_roots_dev(bn254_ctx*, const void* d_roots, size_t n, void* d_coeffs_n_plus_1, void* stream);
int bn254_fr_quotient_coeffs(bn254_ctx*, const void* f_n_plus_1, size_t n, const void* ids, size_t nvec, void* out_scalars);
int bn254_fr_quotient_coeffs_dev(bn254_ctx*, const void* d_f_n_plus_1, size_t n, const void* d_ids, size_t nvec, void* d_out_scalars, void* stream);
int bn254_fr_to_scalars_dev(bn254_ctx*, const void* d_in, size_t n, void* d_out, void* stream);
void bn254_fr_lagrange_basis(const void* s, size_t n, const void* x, void* out);
void bn254_fr_to_scalars(const void* in, size_t n, void* out);

/* diagnostics used by the parity tests: raw Fp Montgomery product, 32 B operands */
int bn254_fp_mul_batch(bn254_ctx*, const void* a, const void* b, size_t n, void* out);

#ifdef __cplusplus
}
#endif
#endif
