package bn254

/*
#include "bn254_b200.h"
*/
import "C"

import (
	"math/big"
	"unsafe"

	gnark "github.com/consensys/gnark-crypto/ecc/bn254"
)

// ---- gnark-named functions (1-element batches; the *Batch functions in batch.go are the throughput path) --------
//
// What one such call costs (B200, host API, against one CPU thread of a C restatement of gnark; INTEGRATION.md has the
// table): Pair 1.3 ms vs 0.8 ms, G1 ScalarMultiplication 1.4 vs 0.17 ms, ScalarMultiplicationBase 0.4 vs 0.17 ms, GT.Exp
// 3.5 vs 0.9 ms.  A lone GPU thread walks ~10^4 dependent 254-bit products; the engine is ahead of sixteen CPU threads from
// about 30 pairings / 130 scalar multiplications per call on.  Code that keeps calling these methods one element at a
// time stays correct and gets slower; hand loops over as slices to the *Batch functions.

// Pair computes prod e(P[i], Q[i]) with one final exponentiation; pairs containing the point at infinity are
// skipped; len(P) == 0 or len(P) != len(Q) returns ErrInvalidSizes (gnark: "invalid inputs sizes").
// Reference call sites: access/tree/access_tree_node.go:106,110, cpabe/bsw07/bsw07_cpabe.go:184,
// ibe/waters05_ibe/waters05_ibe.go:214,259,262, bibe/afp25_bibe/afp25_bibe.go:227,395-403.
func Pair(P []G1Affine, Q []G2Affine) (GT, error) {
	var out GT
	if len(P) == 0 || len(P) != len(Q) {
		return out, ErrInvalidSizes
	}
	err := one(func(d *device) error {
		return d.check(C.bn254_multi_pair_batch(d.ctx, ptr(P), ptr(Q), 1, C.size_t(len(P)), unsafe.Pointer(&out)))
	})
	return out, err
}

// PairingCheck reports Pair(P, Q) == 1 (signature/bls01_signature/bls_signature.go:81-84).
func PairingCheck(P []G1Affine, Q []G2Affine) (bool, error) {
	if len(P) == 0 || len(P) != len(Q) {
		return false, ErrInvalidSizes
	}
	var ok C.uint8_t
	err := one(func(d *device) error {
		return d.check(C.bn254_pairing_check_batch(d.ctx, ptr(P), ptr(Q), 1, C.size_t(len(P)), &ok))
	})
	return ok == 1, err
}

// MillerLoop returns the un-normalised Miller product: defined up to factors FinalExponentiation removes, so only
// FinalExponentiation(MillerLoop(P, Q)) == Pair(P, Q) is canonical (SURVEY.md 8a row 3).
func MillerLoop(P []G1Affine, Q []G2Affine) (GT, error) {
	var out GT
	if len(P) == 0 || len(P) != len(Q) {
		return out, ErrInvalidSizes
	}
	err := one(func(d *device) error {
		return d.check(C.bn254_miller_loop_batch(d.ctx, ptr(P), ptr(Q), 1, C.size_t(len(P)), unsafe.Pointer(&out)))
	})
	return out, err
}

// FinalExponentiation computes (z * _z[0] * ...)^(s (p^12-1)/r) with gnark's cofactor s = 2 x0 (6 x0^2 + 3 x0 + 1).
func FinalExponentiation(z *GT, _z ...*GT) GT {
	acc := *z
	for _, e := range _z {
		acc.Mul(&acc, e)
	}
	var out GT
	if err := one(func(d *device) error {
		return d.check(C.bn254_final_exp_batch(d.ctx, unsafe.Pointer(&acc), 1, unsafe.Pointer(&out)))
	}); err != nil {
		panic(err) // gnark's signature has no error: a lost device cannot be reported any other way
	}
	return out
}

// Generators returns (g1Jac, g2Jac, g1Aff, g2Aff) -- 31 call sites, e.g. signature/bls01_signature/bls_signature.go:32.
func Generators() (g1Jac G1Jac, g2Jac G2Jac, g1Aff G1Affine, g2Aff G2Affine) {
	C.bn254_generators(unsafe.Pointer(&g1Aff), unsafe.Pointer(&g2Aff))
	g1Jac.X, g1Jac.Y = g1Aff.X, g1Aff.Y
	g1Jac.Z.SetOne()
	g2Jac.X, g2Jac.Y = g2Aff.X, g2Aff.Y
	g2Jac.Z.A0.SetOne()
	return
}

// HashToG1 / HashToG2: RFC 9380 hash_to_curve as gnark configures it for BN254 (hash/hash_to.go:113-277 of the
// reference wraps them with its four domain separation tags).
func HashToG1(msg, dst []byte) (G1Affine, error) {
	out, err := HashToG1Batch([][]byte{msg}, dst)
	if err != nil {
		return G1Affine{}, err
	}
	return out[0], nil
}

func HashToG2(msg, dst []byte) (G2Affine, error) {
	out, err := HashToG2Batch([][]byte{msg}, dst)
	if err != nil {
		return G2Affine{}, err
	}
	return out[0], nil
}

// ---- G1Affine methods (receiver is the destination and is returned, as in gnark) ----------------------------------

func (p *G1Affine) Set(a *G1Affine) *G1Affine { *p = *a; return p }
func (p *G1Affine) SetInfinity() *G1Affine    { *p = G1Affine{}; return p }
func (p *G1Affine) IsInfinity() bool          { return *p == G1Affine{} }
func (p *G1Affine) Equal(a *G1Affine) bool    { return *p == *a }
func (p *G1Affine) Neg(a *G1Affine) *G1Affine {
	p.X = a.X
	p.Y.Neg(&a.Y)
	return p
}

// ScalarMultiplication: p = [s]a for any big.Int s (negative: [|s|](-a)); s is reduced mod r.
func (p *G1Affine) ScalarMultiplication(a *G1Affine, s *big.Int) *G1Affine {
	k, neg := scalarBytes(s)
	base := *a
	if neg {
		base.Neg(a)
	}
	mustOne(func(d *device) C.int {
		return C.bn254_g1_mul_batch(d.ctx, unsafe.Pointer(&base), unsafe.Pointer(&k[0]), 1, unsafe.Pointer(p))
	})
	return p
}

// ScalarMultiplicationBase: p = [s]g1.
func (p *G1Affine) ScalarMultiplicationBase(s *big.Int) *G1Affine {
	_, _, g1, _ := Generators()
	return p.ScalarMultiplication(&g1, s)
}

func (p *G1Affine) Add(a, b *G1Affine) *G1Affine {
	x, y := *a, *b
	mustOne(func(d *device) C.int {
		return C.bn254_g1_add_batch(d.ctx, unsafe.Pointer(&x), unsafe.Pointer(&y), 1, unsafe.Pointer(p))
	})
	return p
}

func (p *G1Affine) Sub(a, b *G1Affine) *G1Affine {
	var nb G1Affine
	nb.Neg(b)
	return p.Add(a, &nb)
}

// host-only methods: delegated to gnark on the same bytes
func (p *G1Affine) IsOnCurve() bool                   { return p.gnark().IsOnCurve() }
func (p *G1Affine) IsInSubGroup() bool                { return p.gnark().IsInSubGroup() }
func (p *G1Affine) Bytes() [gnark.SizeOfG1AffineCompressed]byte { return p.gnark().Bytes() }
func (p *G1Affine) RawBytes() [gnark.SizeOfG1AffineUncompressed]byte {
	return p.gnark().RawBytes()
}
func (p *G1Affine) Marshal() []byte                   { return p.gnark().Marshal() }
func (p *G1Affine) Unmarshal(buf []byte) error        { return p.gnark().Unmarshal(buf) }
func (p *G1Affine) SetBytes(buf []byte) (int, error)  { return p.gnark().SetBytes(buf) }
func (p *G1Affine) String() string                    { return p.gnark().String() }

// ---- G2Affine methods -----------------------------------------------------------------------------------------------

func (p *G2Affine) Set(a *G2Affine) *G2Affine { *p = *a; return p }
func (p *G2Affine) SetInfinity() *G2Affine    { *p = G2Affine{}; return p }
func (p *G2Affine) IsInfinity() bool          { return *p == G2Affine{} }
func (p *G2Affine) Equal(a *G2Affine) bool    { return *p == *a }
func (p *G2Affine) Neg(a *G2Affine) *G2Affine {
	p.X = a.X
	p.Y.A0.Neg(&a.Y.A0)
	p.Y.A1.Neg(&a.Y.A1)
	return p
}

func (p *G2Affine) ScalarMultiplication(a *G2Affine, s *big.Int) *G2Affine {
	k, neg := scalarBytes(s)
	base := *a
	if neg {
		base.Neg(a)
	}
	mustOne(func(d *device) C.int {
		return C.bn254_g2_mul_batch(d.ctx, unsafe.Pointer(&base), unsafe.Pointer(&k[0]), 1, unsafe.Pointer(p))
	})
	return p
}

func (p *G2Affine) ScalarMultiplicationBase(s *big.Int) *G2Affine {
	_, _, _, g2 := Generators()
	return p.ScalarMultiplication(&g2, s)
}

func (p *G2Affine) Add(a, b *G2Affine) *G2Affine {
	x, y := *a, *b
	mustOne(func(d *device) C.int {
		return C.bn254_g2_add_batch(d.ctx, unsafe.Pointer(&x), unsafe.Pointer(&y), 1, unsafe.Pointer(p))
	})
	return p
}

func (p *G2Affine) Sub(a, b *G2Affine) *G2Affine {
	var nb G2Affine
	nb.Neg(b)
	return p.Add(a, &nb)
}

func (p *G2Affine) IsOnCurve() bool                   { return p.gnark().IsOnCurve() }
func (p *G2Affine) IsInSubGroup() bool                { return p.gnark().IsInSubGroup() }
func (p *G2Affine) Bytes() [gnark.SizeOfG2AffineCompressed]byte { return p.gnark().Bytes() }
func (p *G2Affine) RawBytes() [gnark.SizeOfG2AffineUncompressed]byte {
	return p.gnark().RawBytes()
}
func (p *G2Affine) Marshal() []byte                   { return p.gnark().Marshal() }
func (p *G2Affine) Unmarshal(buf []byte) error        { return p.gnark().Unmarshal(buf) }
func (p *G2Affine) SetBytes(buf []byte) (int, error)  { return p.gnark().SetBytes(buf) }
func (p *G2Affine) String() string                    { return p.gnark().String() }

// ---- GT methods ---------------------------------------------------------------------------------------------------

func (z *GT) Set(x *GT) *GT    { *z = *x; return z }
func (z *GT) SetOne() *GT      { *z = GT{}; z.C0.B0.A0.SetOne(); return z }
func (z *GT) IsZero() bool     { return *z == GT{} }
func (z *GT) IsOne() bool      { var o GT; o.SetOne(); return *z == o }
func (z *GT) Equal(x *GT) bool { return *z == *x }

// SetRandom delegates to gnark (host randomness; the result is a random Fp12 element, not an element of GT proper --
// the reference's tests only use it that way, e.g. ibe/waters05_ibe/waters05_ibe_test.go).
func (z *GT) SetRandom() (*GT, error) {
	if _, err := z.gnark().SetRandom(); err != nil {
		return nil, err
	}
	return z, nil
}

func (z *GT) Mul(x, y *GT) *GT {
	a, b := *x, *y
	mustOne(func(d *device) C.int {
		return C.bn254_gt_mul_batch(d.ctx, unsafe.Pointer(&a), unsafe.Pointer(&b), 1, unsafe.Pointer(z))
	})
	return z
}

func (z *GT) Div(x, y *GT) *GT {
	a, b := *x, *y
	mustOne(func(d *device) C.int {
		return C.bn254_gt_div_batch(d.ctx, unsafe.Pointer(&a), unsafe.Pointer(&b), 1, unsafe.Pointer(z))
	})
	return z
}

func (z *GT) Inverse(x *GT) *GT {
	var one GT
	one.SetOne()
	return z.Div(&one, x)
}

// Exp: z = x^k, x passed BY VALUE as in gnark (access/tree/access_tree_node.go:156, waters05_ibe.go:219,
// bsw07_cpabe.go:80,146).  k == 0 -> 1; k < 0 -> (x^-1)^|k|.  Generic Fp12 exponentiation (no subgroup assumption).
func (z *GT) Exp(x GT, k *big.Int) *GT {
	e, neg, ok := expBytes(k)
	if !ok { // beyond 256 bits: split k = q 2^256 + r on the host (never happens in the reference: exponents are Fr values)
		q, r := new(big.Int).DivMod(new(big.Int).Abs(k), new(big.Int).Lsh(big.NewInt(1), 256), new(big.Int))
		var hi, lo, t GT
		t.Exp(x, new(big.Int).Lsh(big.NewInt(1), 128))
		t.Exp(t, new(big.Int).Lsh(big.NewInt(1), 128)) // x^(2^256)
		hi.Exp(t, q)
		lo.Exp(x, r)
		z.Mul(&hi, &lo)
		if k.Sign() < 0 {
			z.Inverse(z)
		}
		return z
	}
	base := x
	if neg {
		base.Inverse(&x)
	}
	mustOne(func(d *device) C.int {
		return C.bn254_gt_exp_batch(d.ctx, unsafe.Pointer(&base), unsafe.Pointer(&e[0]), 1, unsafe.Pointer(z))
	})
	return z
}

func (z *GT) Bytes() [gnark.SizeOfGT]byte         { return z.gnark().Bytes() }
func (z *GT) Marshal() []byte                     { return z.gnark().Marshal() }
func (z *GT) Unmarshal(buf []byte) error          { return z.gnark().Unmarshal(buf) }
func (z *GT) SetBytes(buf []byte) error           { return z.gnark().SetBytes(buf) }
func (z *GT) String() string                      { return z.gnark().String() }
func (z *GT) IsInSubGroup() bool                  { return z.gnark().IsInSubGroup() }

// mustOne runs a 1-element call on device 0.  gnark's methods have no error results, so a CUDA failure (lost device,
// out of memory) panics with the library's message instead of returning a silently wrong value.
func mustOne(f func(d *device) C.int) {
	if err := one(func(d *device) error { return d.check(f(d)) }); err != nil {
		panic(err)
	}
}
