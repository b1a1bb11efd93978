package bn254

/*
#include "bn254_b200.h"
*/
import "C"

import (
	"errors"
	"unsafe"
)

// Table handles: immutable precomputations of FIXED operands (public parameters, user keys), replicated on every GPU
// at creation (SURVEY.md 8e) and usable from any goroutine.  Close them before Close().

// G2Lines holds the Miller-loop line coefficients of m fixed G2 points (88 lines x 192 B per point): gnark's
// PrecomputeLines / MillerLoopFixedQ.  For a BSW07 user key (cpabe/bsw07/bsw07_cpabe.go:97-131) the points are
// (Dj..., Dj'..., D); MultiPairLinesBatch then needs no G2 arithmetic at all.
type G2Lines struct {
	m int
	h []*C.bn254_lines
}

func NewG2Lines(Q []G2Affine) (*G2Lines, error) {
	p, err := engines()
	if err != nil {
		return nil, err
	}
	if len(Q) == 0 {
		return nil, ErrInvalidSizes
	}
	t := &G2Lines{m: len(Q), h: make([]*C.bn254_lines, len(p.devs))}
	for i, d := range p.devs {
		if err := d.check(C.bn254_g2_lines_create(d.ctx, ptr(Q), C.size_t(len(Q)), &t.h[i])); err != nil {
			t.Close()
			return nil, err
		}
	}
	return t, nil
}

func (t *G2Lines) Close() {
	for i, h := range t.h {
		if h != nil {
			C.bn254_g2_lines_destroy(h)
			t.h[i] = nil
		}
	}
}

// MultiPairLinesBatch: out[i] = Pair(P[i*m:(i+1)*m], Q) for the m table points -- bit-identical to MultiPairBatch on
// the same operands.
func MultiPairLinesBatch(P []G1Affine, t *G2Lines) ([]GT, error) {
	if t == nil || len(P)%t.m != 0 {
		return nil, ErrInvalidSizes
	}
	m := t.m
	n := len(P) / m
	out := make([]GT, n)
	err := shard(n, wave/m+1, func(d *device, lo, hi int) error {
		return d.check(C.bn254_multi_pair_lines_batch(d.ctx, ptr(P[lo*m:hi*m]), t.h[d.ordinal], C.size_t(hi-lo), ptr(out[lo:hi])))
	})
	return out, err
}

// FixedBase is a 32 x 255 window table of one base: group 1 = G1, 2 = G2, 3 = GT.
type FixedBase struct {
	group int
	h     []*C.bn254_fixed_base
}

func newFixedBase(group int, base unsafe.Pointer) (*FixedBase, error) {
	p, err := engines()
	if err != nil {
		return nil, err
	}
	t := &FixedBase{group: group, h: make([]*C.bn254_fixed_base, len(p.devs))}
	for i, d := range p.devs {
		if err := d.check(C.bn254_fixed_base_create(d.ctx, C.int(group), base, &t.h[i])); err != nil {
			t.Close()
			return nil, err
		}
	}
	return t, nil
}

func NewFixedBaseG1(base *G1Affine) (*FixedBase, error) { b := *base; return newFixedBase(C.BN254_GROUP_G1, unsafe.Pointer(&b)) }
func NewFixedBaseG2(base *G2Affine) (*FixedBase, error) { b := *base; return newFixedBase(C.BN254_GROUP_G2, unsafe.Pointer(&b)) }
func NewFixedBaseGT(base *GT) (*FixedBase, error)       { b := *base; return newFixedBase(C.BN254_GROUP_GT, unsafe.Pointer(&b)) }

func (t *FixedBase) Close() {
	for i, h := range t.h {
		if h != nil {
			C.bn254_fixed_base_destroy(h)
			t.h[i] = nil
		}
	}
}

var errWrongGroup = errors.New("bn254: fixed-base table of another group")

func (t *FixedBase) MulBatchG1(s []Scalar) ([]G1Affine, error) {
	if t.group != C.BN254_GROUP_G1 {
		return nil, errWrongGroup
	}
	out := make([]G1Affine, len(s))
	err := shard(len(s), wave, func(d *device, lo, hi int) error {
		return d.check(C.bn254_g1_fixed_mul_batch(d.ctx, t.h[d.ordinal], ptr(s[lo:hi]), C.size_t(hi-lo), ptr(out[lo:hi])))
	})
	return out, err
}

func (t *FixedBase) MulBatchG2(s []Scalar) ([]G2Affine, error) {
	if t.group != C.BN254_GROUP_G2 {
		return nil, errWrongGroup
	}
	out := make([]G2Affine, len(s))
	err := shard(len(s), wave, func(d *device, lo, hi int) error {
		return d.check(C.bn254_g2_fixed_mul_batch(d.ctx, t.h[d.ordinal], ptr(s[lo:hi]), C.size_t(hi-lo), ptr(out[lo:hi])))
	})
	return out, err
}

func (t *FixedBase) ExpBatchGT(k []Scalar) ([]GT, error) {
	if t.group != C.BN254_GROUP_GT {
		return nil, errWrongGroup
	}
	out := make([]GT, len(k))
	err := shard(len(k), wave, func(d *device, lo, hi int) error {
		return d.check(C.bn254_gt_fixed_exp_batch(d.ctx, t.h[d.ordinal], ptr(k[lo:hi]), C.size_t(hi-lo), ptr(out[lo:hi])))
	})
	return out, err
}

// MSMTable holds per-point window tables of `len` shared points: out[v] = sum_j [s[v*len + j]] P_j for many
// coefficient vectors over the same points (bibe/afp25_bibe/afp25_bibe_utils.go:45-55 computeG1PolynomialTau over
// g1, [tau]1 .. [tau^B]1; bibe/gwww25_bibe/gwww25_bibe_utils.go:40-50 on G2).
type MSMTable struct {
	group, length int
	h             []*C.bn254_msm_table
}

func newMSMTable(group int, points unsafe.Pointer, n int) (*MSMTable, error) {
	p, err := engines()
	if err != nil {
		return nil, err
	}
	if n == 0 {
		return nil, ErrInvalidSizes
	}
	t := &MSMTable{group: group, length: n, h: make([]*C.bn254_msm_table, len(p.devs))}
	for i, d := range p.devs {
		if err := d.check(C.bn254_msm_table_create(d.ctx, C.int(group), points, C.size_t(n), &t.h[i])); err != nil {
			t.Close()
			return nil, err
		}
	}
	return t, nil
}

func NewMSMTableG1(points []G1Affine) (*MSMTable, error) { return newMSMTable(C.BN254_GROUP_G1, ptr(points), len(points)) }
func NewMSMTableG2(points []G2Affine) (*MSMTable, error) { return newMSMTable(C.BN254_GROUP_G2, ptr(points), len(points)) }

func (t *MSMTable) Close() {
	for i, h := range t.h {
		if h != nil {
			C.bn254_msm_table_destroy(h)
			t.h[i] = nil
		}
	}
}

// Len is the number of points (= scalars per vector).
func (t *MSMTable) Len() int { return t.length }

func (t *MSMTable) BatchG1(s []Scalar) ([]G1Affine, error) {
	if t.group != C.BN254_GROUP_G1 || len(s)%t.length != 0 {
		return nil, ErrInvalidSizes
	}
	n, l := len(s)/t.length, t.length
	out := make([]G1Affine, n)
	err := shard(n, wave/l+1, func(d *device, lo, hi int) error {
		return d.check(C.bn254_msm_batch(d.ctx, t.h[d.ordinal], ptr(s[lo*l:hi*l]), C.size_t(hi-lo), ptr(out[lo:hi])))
	})
	return out, err
}

func (t *MSMTable) BatchG2(s []Scalar) ([]G2Affine, error) {
	if t.group != C.BN254_GROUP_G2 || len(s)%t.length != 0 {
		return nil, ErrInvalidSizes
	}
	n, l := len(s)/t.length, t.length
	out := make([]G2Affine, n)
	err := shard(n, wave/l+1, func(d *device, lo, hi int) error {
		return d.check(C.bn254_msm_batch(d.ctx, t.h[d.ordinal], ptr(s[lo*l:hi*l]), C.size_t(hi-lo), ptr(out[lo:hi])))
	})
	return out, err
}
