// Package bn254 mirrors the part of github.com/consensys/gnark-crypto/ecc/bn254 (v0.19.0) that the reference
// schemes call -- same names, same signatures, same error values -- and routes the arithmetic to the B200 engine
// through the C ABI of include/bn254_b200.h.  A scheme switches over by changing two imports:
//
//	"github.com/consensys/gnark-crypto/ecc/bn254"     -> "github.com/mmsyan/GoPairingBasedCryptography/b200/bn254"
//	"github.com/consensys/gnark-crypto/ecc/bn254/fr"  -> "github.com/mmsyan/GoPairingBasedCryptography/b200/bn254/fr"
//
// The value types are this package's OWN (not aliases): Go resolves methods by the receiver's defining package, so
// only own types let `z.ScalarMultiplication(&a, s)` (119 call sites, e.g. signature/bls01_signature/bls_signature.go:63)
// and `z.Exp(x, k)` (41 call sites, e.g. access/tree/access_tree_node.go:156) reach the GPU.  They have gnark's exact
// memory layout (arrays of uint64 limbs in Montgomery form), so they stay comparable with ==, usable as map keys,
// zero-value meaningful (zero GT is 0, zero affine point is infinity), and convert to gnark's types with an
// unsafe cast where a host-only method (Bytes, Marshal, String, IsInSubGroup) is delegated.
package bn254

import (
	"math/big"
	"unsafe"

	gnark "github.com/consensys/gnark-crypto/ecc/bn254"
	"github.com/consensys/gnark-crypto/ecc/bn254/fp"
	gfr "github.com/consensys/gnark-crypto/ecc/bn254/fr"
)

// E2, E6 mirror gnark's internal/fptower.E2 / E6 field order (SURVEY.md 8c item 1).
type E2 struct{ A0, A1 fp.Element }
type E6 struct{ B0, B1, B2 E2 }

// G1Affine is a point of G1 in affine coordinates; (0, 0) is the point at infinity.  64 bytes.
type G1Affine struct{ X, Y fp.Element }

// G2Affine is a point of G2 (on the twist) in affine coordinates; (0, 0) is the point at infinity.  128 bytes.
type G2Affine struct{ X, Y E2 }

// G1Jac / G2Jac exist because Generators() returns them; the reference never computes with them.
type G1Jac struct{ X, Y, Z fp.Element }
type G2Jac struct{ X, Y, Z E2 }

// GT is an element of Fp12 = E12{C0, C1 E6}.  384 bytes.  The zero value is 0, not 1.
type GT struct{ C0, C1 E6 }

// compile-time layout checks against gnark's types: the unsafe casts below are sound only while these hold
var (
	_ [unsafe.Sizeof(gnark.G1Affine{})]byte = [unsafe.Sizeof(G1Affine{})]byte{}
	_ [unsafe.Sizeof(gnark.G2Affine{})]byte = [unsafe.Sizeof(G2Affine{})]byte{}
	_ [unsafe.Sizeof(gnark.GT{})]byte       = [unsafe.Sizeof(GT{})]byte{}
	_ [64]byte                              = [unsafe.Sizeof(G1Affine{})]byte{}
	_ [128]byte                             = [unsafe.Sizeof(G2Affine{})]byte{}
	_ [384]byte                             = [unsafe.Sizeof(GT{})]byte{}
	_ [32]byte                              = [unsafe.Sizeof(gfr.Element{})]byte{}
)

func (p *G1Affine) gnark() *gnark.G1Affine { return (*gnark.G1Affine)(unsafe.Pointer(p)) }
func (p *G2Affine) gnark() *gnark.G2Affine { return (*gnark.G2Affine)(unsafe.Pointer(p)) }
func (z *GT) gnark() *gnark.GT             { return (*gnark.GT)(unsafe.Pointer(z)) }

// FromGnarkG1 etc. convert values that crossed from code still on gnark's package (no copy of semantics: same bytes).
func FromGnarkG1(p *gnark.G1Affine) G1Affine { return *(*G1Affine)(unsafe.Pointer(p)) }
func FromGnarkG2(p *gnark.G2Affine) G2Affine { return *(*G2Affine)(unsafe.Pointer(p)) }
func FromGnarkGT(z *gnark.GT) GT             { return *(*GT)(unsafe.Pointer(z)) }

// scalarBytes converts a big.Int exactly as the engine's ABI wants it: 32 little-endian bytes of the value reduced
// into [0, r).  neg reports s < 0 (gnark negates the point / inverts the GT element and uses |s|).
func scalarBytes(s *big.Int) (out [32]byte, neg bool) {
	v := s
	if s.Sign() < 0 {
		neg = true
		v = new(big.Int).Neg(s)
	}
	if v.BitLen() > 254 || v.Cmp(gfr.Modulus()) >= 0 {
		v = new(big.Int).Mod(v, gfr.Modulus())
	}
	be := v.Bytes() // big-endian, no leading zeros
	for i, b := range be {
		out[len(be)-1-i] = b
	}
	return
}

// expBytes is scalarBytes for GT.Exp: gnark's E12.Exp does NOT reduce the exponent mod r (the element need not lie
// in GT proper); exponents of up to 256 bits go to the engine as they are, longer ones are rejected by the caller.
func expBytes(k *big.Int) (out [32]byte, neg bool, ok bool) {
	v := k
	if k.Sign() < 0 {
		neg = true
		v = new(big.Int).Neg(k)
	}
	if v.BitLen() > 256 {
		return out, neg, false
	}
	be := v.Bytes()
	for i, b := range be {
		out[len(be)-1-i] = b
	}
	return out, neg, true
}
