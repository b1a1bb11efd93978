package bn254

/*
#include "bn254_b200.h"
*/
import "C"

import (
	"math/big"
	"unsafe"

	gfr "github.com/consensys/gnark-crypto/ecc/bn254/fr"
)

// Batch entry points: what actually reaches the GPU at throughput.  Every function splits its batch into
// ceil(n/G) contiguous chunks over the G visible GPUs (no inter-GPU traffic) and returns when all results are in
// the output slice.  Slices may live in PinnedBytes memory to skip the staging copy.

// wave is the smallest chunk worth a GPU of its own: one full wave of the one-thread-per-element kernels.
const wave = 148 * 3 * 128

// Scalar is the engine's scalar format: the big.Int value as 32 little-endian bytes.
type Scalar [32]byte

// ScalarFromBig reduces s into [0, r) (callers negate the point for negative s, as gnark does).
func ScalarFromBig(s *big.Int) Scalar { k, _ := scalarBytes(new(big.Int).Abs(s)); return Scalar(k) }

// ScalarsFromFr is x.BigInt(new(big.Int)) for a whole slice, without the big.Int detour: Montgomery -> regular form.
func ScalarsFromFr(x []gfr.Element) []Scalar {
	out := make([]Scalar, len(x))
	if len(x) > 0 {
		C.bn254_fr_to_scalars(ptr(x), C.size_t(len(x)), ptr(out))
	}
	return out
}

// PairBatch: out[i] = Pair({P[i]}, {Q[i]}) -- BASELINE configs[1].
func PairBatch(P []G1Affine, Q []G2Affine) ([]GT, error) {
	if len(P) != len(Q) {
		return nil, ErrInvalidSizes
	}
	out := make([]GT, len(P))
	err := shard(len(P), wave, func(d *device, lo, hi int) error {
		return d.check(C.bn254_pair_batch(d.ctx, ptr(P[lo:hi]), ptr(Q[lo:hi]), C.size_t(hi-lo), ptr(out[lo:hi])))
	})
	return out, err
}

// MultiPairBatch: out[i] = Pair(P[i*k:(i+1)*k], Q[i*k:(i+1)*k]) -- one Miller product and ONE final exponentiation
// per row (the fused decryption shape of access/tree/access_tree_node.go:96-164 + cpabe/bsw07/bsw07_cpabe.go:172-195).
func MultiPairBatch(P []G1Affine, Q []G2Affine, k int) ([]GT, error) {
	if k <= 0 || len(P) != len(Q) || len(P)%k != 0 {
		return nil, ErrInvalidSizes
	}
	n := len(P) / k
	out := make([]GT, n)
	err := shard(n, wave/k+1, func(d *device, lo, hi int) error {
		return d.check(C.bn254_multi_pair_batch(d.ctx, ptr(P[lo*k:hi*k]), ptr(Q[lo*k:hi*k]), C.size_t(hi-lo), C.size_t(k), ptr(out[lo:hi])))
	})
	return out, err
}

// MillerLoopBatch / FinalExponentiationBatch: the two halves of MultiPairBatch.
func MillerLoopBatch(P []G1Affine, Q []G2Affine, k int) ([]GT, error) {
	if k <= 0 || len(P) != len(Q) || len(P)%k != 0 {
		return nil, ErrInvalidSizes
	}
	n := len(P) / k
	out := make([]GT, n)
	err := shard(n, wave/k+1, func(d *device, lo, hi int) error {
		return d.check(C.bn254_miller_loop_batch(d.ctx, ptr(P[lo*k:hi*k]), ptr(Q[lo*k:hi*k]), C.size_t(hi-lo), C.size_t(k), ptr(out[lo:hi])))
	})
	return out, err
}

func FinalExponentiationBatch(z []GT) ([]GT, error) {
	out := make([]GT, len(z))
	err := shard(len(z), wave, func(d *device, lo, hi int) error {
		return d.check(C.bn254_final_exp_batch(d.ctx, ptr(z[lo:hi]), C.size_t(hi-lo), ptr(out[lo:hi])))
	})
	return out, err
}

// PairingCheckBatch: ok[i] = PairingCheck(P[i*k:(i+1)*k], Q[i*k:(i+1)*k]).
func PairingCheckBatch(P []G1Affine, Q []G2Affine, k int) ([]bool, error) {
	if k <= 0 || len(P) != len(Q) || len(P)%k != 0 {
		return nil, ErrInvalidSizes
	}
	n := len(P) / k
	raw := make([]byte, n)
	err := shard(n, wave/k+1, func(d *device, lo, hi int) error {
		return d.check(C.bn254_pairing_check_batch(d.ctx, ptr(P[lo*k:hi*k]), ptr(Q[lo*k:hi*k]), C.size_t(hi-lo), C.size_t(k), (*C.uint8_t)(ptr(raw[lo:hi]))))
	})
	return bools(raw), err
}

// BLSVerifyBatch: ok[i] = PairingCheck({pk, -g1}, {hm[i], sigma[i]}) -- signature/bls01_signature/bls_signature.go:71-89
// for one public key and n (message hash, signature) pairs; the two G1 points are sent once.
func BLSVerifyBatch(pk *G1Affine, hm, sigma []G2Affine) ([]bool, error) {
	if len(hm) != len(sigma) {
		return nil, ErrInvalidSizes
	}
	_, _, g1, _ := Generators()
	var p01 [2]G1Affine
	p01[0] = *pk
	p01[1].Neg(&g1)
	raw := make([]byte, len(hm))
	err := shard(len(hm), wave, func(d *device, lo, hi int) error {
		return d.check(C.bn254_pairing_check2_fixed_g1_batch(d.ctx, unsafe.Pointer(&p01[0]), ptr(hm[lo:hi]), ptr(sigma[lo:hi]), C.size_t(hi-lo), (*C.uint8_t)(ptr(raw[lo:hi]))))
	})
	return bools(raw), err
}

func bools(raw []byte) []bool {
	out := make([]bool, len(raw))
	for i, b := range raw {
		out[i] = b != 0
	}
	return out
}

// ScalarMulBatchG1 / G2: out[i] = [s[i]] base[i]  (variable base, 2-dimensional GLV).
func ScalarMulBatchG1(base []G1Affine, s []Scalar) ([]G1Affine, error) {
	if len(base) != len(s) {
		return nil, ErrInvalidSizes
	}
	out := make([]G1Affine, len(s))
	err := shard(len(s), wave, func(d *device, lo, hi int) error {
		return d.check(C.bn254_g1_mul_batch(d.ctx, ptr(base[lo:hi]), ptr(s[lo:hi]), C.size_t(hi-lo), ptr(out[lo:hi])))
	})
	return out, err
}

func ScalarMulBatchG2(base []G2Affine, s []Scalar) ([]G2Affine, error) {
	if len(base) != len(s) {
		return nil, ErrInvalidSizes
	}
	out := make([]G2Affine, len(s))
	err := shard(len(s), wave, func(d *device, lo, hi int) error {
		return d.check(C.bn254_g2_mul_batch(d.ctx, ptr(base[lo:hi]), ptr(s[lo:hi]), C.size_t(hi-lo), ptr(out[lo:hi])))
	})
	return out, err
}

// FixedBaseBatchG1 / G2: out[i] = [s[i]] base for ONE base (ScalarMultiplicationBase and every other public-parameter
// base).  Each GPU keeps window tables of recently used bases; see FixedBase for an explicit handle.
func FixedBaseBatchG1(base *G1Affine, s []Scalar) ([]G1Affine, error) {
	out := make([]G1Affine, len(s))
	b := *base
	err := shard(len(s), wave, func(d *device, lo, hi int) error {
		return d.check(C.bn254_g1_mul_base_batch(d.ctx, unsafe.Pointer(&b), ptr(s[lo:hi]), C.size_t(hi-lo), ptr(out[lo:hi])))
	})
	return out, err
}

func FixedBaseBatchG2(base *G2Affine, s []Scalar) ([]G2Affine, error) {
	out := make([]G2Affine, len(s))
	b := *base
	err := shard(len(s), wave, func(d *device, lo, hi int) error {
		return d.check(C.bn254_g2_mul_base_batch(d.ctx, unsafe.Pointer(&b), ptr(s[lo:hi]), C.size_t(hi-lo), ptr(out[lo:hi])))
	})
	return out, err
}

// AddBatchG1 / G2: out[i] = a[i] + b[i] with gnark's Add semantics (infinity, doubling, P + (-P)).
func AddBatchG1(a, b []G1Affine) ([]G1Affine, error) {
	if len(a) != len(b) {
		return nil, ErrInvalidSizes
	}
	out := make([]G1Affine, len(a))
	err := shard(len(a), wave, func(d *device, lo, hi int) error {
		return d.check(C.bn254_g1_add_batch(d.ctx, ptr(a[lo:hi]), ptr(b[lo:hi]), C.size_t(hi-lo), ptr(out[lo:hi])))
	})
	return out, err
}

func AddBatchG2(a, b []G2Affine) ([]G2Affine, error) {
	if len(a) != len(b) {
		return nil, ErrInvalidSizes
	}
	out := make([]G2Affine, len(a))
	err := shard(len(a), wave, func(d *device, lo, hi int) error {
		return d.check(C.bn254_g2_add_batch(d.ctx, ptr(a[lo:hi]), ptr(b[lo:hi]), C.size_t(hi-lo), ptr(out[lo:hi])))
	})
	return out, err
}

// SumBatchG1 / G2: out[g] = points[g*length] + ... + points[(g+1)*length-1]  (the Add chains of
// bibe/afp25_bibe/afp25_bibe_utils.go:45-55 and gwww25's G2-side MSM; one inversion per partial sum, not per Add).
func SumBatchG1(points []G1Affine, length int) ([]G1Affine, error) {
	if length <= 0 || len(points)%length != 0 {
		return nil, ErrInvalidSizes
	}
	n := len(points) / length
	out := make([]G1Affine, n)
	err := shard(n, wave/length+1, func(d *device, lo, hi int) error {
		return d.check(C.bn254_g1_sum_batch(d.ctx, ptr(points[lo*length:hi*length]), C.size_t(hi-lo), C.size_t(length), ptr(out[lo:hi])))
	})
	return out, err
}

func SumBatchG2(points []G2Affine, length int) ([]G2Affine, error) {
	if length <= 0 || len(points)%length != 0 {
		return nil, ErrInvalidSizes
	}
	n := len(points) / length
	out := make([]G2Affine, n)
	err := shard(n, wave/length+1, func(d *device, lo, hi int) error {
		return d.check(C.bn254_g2_sum_batch(d.ctx, ptr(points[lo*length:hi*length]), C.size_t(hi-lo), C.size_t(length), ptr(out[lo:hi])))
	})
	return out, err
}

// SubsetSumBatchG2: out[i] = U[0] + sum_{j: bit j of sel[i]} U[j+1], bit j = bit (7 - j%8) of byte j/8 -- the Waters
// hash of ibe/waters05_ibe/waters05_ibe.go:227-233 with the identity-vector order of :302-313.  len(U) = m + 1,
// len(sel) = n * ceil(m/8).
func SubsetSumBatchG2(U []G2Affine, sel []byte) ([]G2Affine, error) {
	m := len(U) - 1
	row := (m + 7) / 8
	if m <= 0 || len(sel)%row != 0 {
		return nil, ErrInvalidSizes
	}
	n := len(sel) / row
	out := make([]G2Affine, n)
	err := shard(n, wave, func(d *device, lo, hi int) error {
		return d.check(C.bn254_g2_subset_sum_batch(d.ctx, ptr(U), C.size_t(m), ptr(sel[lo*row:hi*row]), C.size_t(hi-lo), ptr(out[lo:hi])))
	})
	return out, err
}

func SubsetSumBatchG1(U []G1Affine, sel []byte) ([]G1Affine, error) {
	m := len(U) - 1
	row := (m + 7) / 8
	if m <= 0 || len(sel)%row != 0 {
		return nil, ErrInvalidSizes
	}
	n := len(sel) / row
	out := make([]G1Affine, n)
	err := shard(n, wave, func(d *device, lo, hi int) error {
		return d.check(C.bn254_g1_subset_sum_batch(d.ctx, ptr(U), C.size_t(m), ptr(sel[lo*row:hi*row]), C.size_t(hi-lo), ptr(out[lo:hi])))
	})
	return out, err
}

// GTExpBatch: out[i] = x[i]^k[i], generic Fp12 (gnark's E12.Exp).  GTExpCyclotomicBatch: same result for x in GT
// proper (every Pair output and products / quotients / powers of them -- all GT.Exp bases of the reference), ~2.5x
// less work (GLV split of the exponent, Granger-Scott squarings); undefined for other Fp12 elements.
func GTExpBatch(x []GT, k []Scalar) ([]GT, error) { return gtExp(x, k, false) }

func GTExpCyclotomicBatch(x []GT, k []Scalar) ([]GT, error) { return gtExp(x, k, true) }

func gtExp(x []GT, k []Scalar, cyclo bool) ([]GT, error) {
	if len(x) != len(k) {
		return nil, ErrInvalidSizes
	}
	out := make([]GT, len(x))
	err := shard(len(x), wave, func(d *device, lo, hi int) error {
		if cyclo {
			return d.check(C.bn254_gt_cyclo_exp_batch(d.ctx, ptr(x[lo:hi]), ptr(k[lo:hi]), C.size_t(hi-lo), ptr(out[lo:hi])))
		}
		return d.check(C.bn254_gt_exp_batch(d.ctx, ptr(x[lo:hi]), ptr(k[lo:hi]), C.size_t(hi-lo), ptr(out[lo:hi])))
	})
	return out, err
}

// GTExpBaseBatch: out[i] = x^k[i] for ONE base (waters05_ibe.go:219: e(g1,g2)^alpha raised to every t).
func GTExpBaseBatch(x *GT, k []Scalar) ([]GT, error) {
	out := make([]GT, len(k))
	b := *x
	err := shard(len(k), wave, func(d *device, lo, hi int) error {
		return d.check(C.bn254_gt_exp_base_batch(d.ctx, unsafe.Pointer(&b), ptr(k[lo:hi]), C.size_t(hi-lo), ptr(out[lo:hi])))
	})
	return out, err
}

// GTMulBatch / GTDivBatch: out[i] = a[i] * b[i], a[i] / b[i].  GTDivCyclotomicBatch: the same quotient for divisors in
// GT proper (pairing outputs and their products / powers -- the divisor of every Div in the reference's decryption
// flows): b^-1 = conj(b), so it is one Fp12 product instead of an inversion and a product.
func GTMulBatch(a, b []GT) ([]GT, error)           { return gtBinary(a, b, 0) }
func GTDivBatch(a, b []GT) ([]GT, error)           { return gtBinary(a, b, 1) }
func GTDivCyclotomicBatch(a, b []GT) ([]GT, error) { return gtBinary(a, b, 2) }

func gtBinary(a, b []GT, mode int) ([]GT, error) {
	if len(a) != len(b) {
		return nil, ErrInvalidSizes
	}
	out := make([]GT, len(a))
	err := shard(len(a), wave, func(d *device, lo, hi int) error {
		switch mode {
		case 1:
			return d.check(C.bn254_gt_div_batch(d.ctx, ptr(a[lo:hi]), ptr(b[lo:hi]), C.size_t(hi-lo), ptr(out[lo:hi])))
		case 2:
			return d.check(C.bn254_gt_cyclo_div_batch(d.ctx, ptr(a[lo:hi]), ptr(b[lo:hi]), C.size_t(hi-lo), ptr(out[lo:hi])))
		}
		return d.check(C.bn254_gt_mul_batch(d.ctx, ptr(a[lo:hi]), ptr(b[lo:hi]), C.size_t(hi-lo), ptr(out[lo:hi])))
	})
	return out, err
}

// HashToG1Batch / HashToG2Batch: n x bn254.HashToG1(msg, dst) / HashToG2 on the GPU (SHA-256 expand_message_xmd, SVDW
// map, cofactor clearing).  len(dst) > 255 is an error, as in gnark.
func HashToG1Batch(msgs [][]byte, dst []byte) ([]G1Affine, error) {
	out := make([]G1Affine, len(msgs))
	err := hashBatch(msgs, dst, func(d *device, blob []byte, off []uint64, lo, hi int) error {
		return d.check(C.bn254_hash_to_g1_batch(d.ctx, (*C.uint8_t)(ptr(blob)), (*C.uint64_t)(ptr(off[lo:hi+1])), C.size_t(hi-lo),
			(*C.uint8_t)(ptr(dst)), C.size_t(len(dst)), ptr(out[lo:hi])))
	})
	return out, err
}

func HashToG2Batch(msgs [][]byte, dst []byte) ([]G2Affine, error) {
	out := make([]G2Affine, len(msgs))
	err := hashBatch(msgs, dst, func(d *device, blob []byte, off []uint64, lo, hi int) error {
		return d.check(C.bn254_hash_to_g2_batch(d.ctx, (*C.uint8_t)(ptr(blob)), (*C.uint64_t)(ptr(off[lo:hi+1])), C.size_t(hi-lo),
			(*C.uint8_t)(ptr(dst)), C.size_t(len(dst)), ptr(out[lo:hi])))
	})
	return out, err
}

func hashBatch(msgs [][]byte, dst []byte, f func(d *device, blob []byte, off []uint64, lo, hi int) error) error {
	off := make([]uint64, len(msgs)+1)
	total := 0
	for i, m := range msgs {
		total += len(m)
		off[i+1] = uint64(total)
	}
	blob := make([]byte, total+1) // never empty: the C side wants a non-NULL pointer
	p := 0
	for _, m := range msgs {
		p += copy(blob[p:], m)
	}
	// offsets are absolute into blob; each shard passes its slice of them and the library rebases on offsets[0]
	return shard(len(msgs), wave, func(d *device, lo, hi int) error { return f(d, blob, off, lo, hi) })
}
