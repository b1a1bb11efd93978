package bn254_test

// Parity against gnark-crypto itself: run with `go test ./...` on a machine with a B200, the built
// libbn254_b200.so and a Go toolchain (this repository's build image has none; see INTEGRATION.md).  Every test feeds
// the same inputs to gnark and to this package and compares the results bit for bit.

import (
	"crypto/rand"
	"math/big"
	"testing"
	"unsafe"

	gnark "github.com/consensys/gnark-crypto/ecc/bn254"
	gfr "github.com/consensys/gnark-crypto/ecc/bn254/fr"

	b200 "github.com/mmsyan/GoPairingBasedCryptography/b200/bn254"
)

func randScalar(t *testing.T) *big.Int {
	var e gfr.Element
	if _, err := e.SetRandom(); err != nil {
		t.Fatal(err)
	}
	return e.BigInt(new(big.Int))
}

func gG1(p *b200.G1Affine) *gnark.G1Affine { return (*gnark.G1Affine)(unsafe.Pointer(p)) }
func gG2(p *b200.G2Affine) *gnark.G2Affine { return (*gnark.G2Affine)(unsafe.Pointer(p)) }
func gGT(p *b200.GT) *gnark.GT             { return (*gnark.GT)(unsafe.Pointer(p)) }

func TestGeneratorsMatchGnark(t *testing.T) {
	_, _, g1, g2 := b200.Generators()
	_, _, h1, h2 := gnark.Generators()
	if *gG1(&g1) != h1 || *gG2(&g2) != h2 {
		t.Fatal("generators differ")
	}
}

func TestScalarMultiplicationAndAddMatchGnark(t *testing.T) {
	_, _, g1, g2 := b200.Generators()
	for i := 0; i < 8; i++ {
		s, u := randScalar(t), randScalar(t)
		if i == 0 {
			s = new(big.Int).Neg(s) // negative scalars negate the point
		}
		var a, b, c b200.G1Affine
		a.ScalarMultiplication(&g1, s)
		b.ScalarMultiplicationBase(u)
		c.Add(&a, &b)
		var ga, gb, gc gnark.G1Affine
		ga.ScalarMultiplication(gG1(&g1), s)
		gb.ScalarMultiplicationBase(u)
		gc.Add(&ga, &gb)
		if *gG1(&a) != ga || *gG1(&b) != gb || *gG1(&c) != gc {
			t.Fatalf("G1 mismatch at %d", i)
		}
		var x, y, z b200.G2Affine
		x.ScalarMultiplication(&g2, s)
		y.ScalarMultiplicationBase(u)
		z.Sub(&x, &y)
		var gx, gy, gz gnark.G2Affine
		gx.ScalarMultiplication(gG2(&g2), s)
		gy.ScalarMultiplicationBase(u)
		gz.Sub(&gx, &gy)
		if *gG2(&x) != gx || *gG2(&y) != gy || *gG2(&z) != gz {
			t.Fatalf("G2 mismatch at %d", i)
		}
	}
}

func TestPairAndGTMatchGnark(t *testing.T) {
	_, _, g1, g2 := b200.Generators()
	var P [3]b200.G1Affine
	var Q [3]b200.G2Affine
	for i := range P {
		P[i].ScalarMultiplication(&g1, randScalar(t))
		Q[i].ScalarMultiplication(&g2, randScalar(t))
	}
	gP := []gnark.G1Affine{*gG1(&P[0]), *gG1(&P[1]), *gG1(&P[2])}
	gQ := []gnark.G2Affine{*gG2(&Q[0]), *gG2(&Q[1]), *gG2(&Q[2])}
	for k := 1; k <= 3; k++ {
		e, err := b200.Pair(P[:k], Q[:k])
		if err != nil {
			t.Fatal(err)
		}
		ge, _ := gnark.Pair(gP[:k], gQ[:k])
		if *gGT(&e) != ge {
			t.Fatalf("Pair(k=%d): GT bytes differ from gnark (final-exponent cofactor?)", k)
		}
		if e.Bytes() != ge.Bytes() {
			t.Fatal("GT.Bytes differ")
		}
	}
	if _, err := b200.Pair(nil, nil); err == nil || err.Error() != "invalid inputs sizes" {
		t.Fatal("empty Pair must return gnark's error")
	}
	e, _ := b200.Pair(P[:1], Q[:1])
	ge, _ := gnark.Pair(gP[:1], gQ[:1])
	k := randScalar(t)
	var x, y, inv b200.GT
	x.Exp(e, k)
	y.Exp(e, new(big.Int).Neg(k))
	inv.Inverse(&x)
	var gx gnark.GT
	gx.Exp(ge, k)
	if *gGT(&x) != gx || y != inv {
		t.Fatal("GT.Exp differs from gnark")
	}
	ml, _ := b200.MillerLoop(P[:2], Q[:2])
	fe := b200.FinalExponentiation(&ml)
	e2, _ := b200.Pair(P[:2], Q[:2])
	if fe != e2 {
		t.Fatal("FinalExponentiation(MillerLoop) != Pair")
	}
	ok, _ := b200.PairingCheck(P[:2], Q[:2])
	gok, _ := gnark.PairingCheck(gP[:2], gQ[:2])
	if ok != gok {
		t.Fatal("PairingCheck differs")
	}
}

func TestHashToCurveMatchesGnark(t *testing.T) {
	dst := []byte("Hash Bytes To Element In G2") // hash/hash_to.go:272
	for _, m := range [][]byte{nil, []byte("abc"), make([]byte, 200)} {
		a, err := b200.HashToG2(m, dst)
		if err != nil {
			t.Fatal(err)
		}
		b, _ := gnark.HashToG2(m, dst)
		if *gG2(&a) != b {
			t.Fatal("HashToG2 differs from gnark")
		}
		c, _ := b200.HashToG1(m, dst)
		d, _ := gnark.HashToG1(m, dst)
		if *gG1(&c) != d {
			t.Fatal("HashToG1 differs from gnark")
		}
	}
}

func TestBatchEntryPoints(t *testing.T) {
	_, _, g1, g2 := b200.Generators()
	n := 1000
	s := make([]b200.Scalar, n)
	for i := range s {
		rand.Read(s[i][:31])
	}
	P, err := b200.FixedBaseBatchG1(&g1, s)
	if err != nil {
		t.Fatal(err)
	}
	Q, _ := b200.FixedBaseBatchG2(&g2, s)
	E, err := b200.PairBatch(P, Q)
	if err != nil {
		t.Fatal(err)
	}
	for _, i := range []int{0, 1, n / 2, n - 1} {
		ge, _ := gnark.Pair([]gnark.G1Affine{*gG1(&P[i])}, []gnark.G2Affine{*gG2(&Q[i])})
		if *gGT(&E[i]) != ge {
			t.Fatalf("PairBatch[%d] differs from gnark", i)
		}
	}
}
