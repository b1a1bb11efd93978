// gen_gnark_fixtures -- emits tests/golden/gnark_fixtures.json: outputs of gnark-crypto v0.19.0 `ecc/bn254` on
// deterministic inputs, in gnark's in-memory layout (Montgomery limbs, little-endian bytes = what the C ABI takes)
// and in its wire encodings.  tests/test_gnark_fixtures.py compares the oracle (CPU) and the CUDA engine (GPU)
// against this file when it exists.  It closes the "parity unpinned" items of DESIGN.md section 2: the GT bytes of
// Pair (final-exponent cofactor), G2 ScalarMultiplication, HashToG2 (Z, cofactor multiple), GT.Bytes / Marshal.
//
// Test infrastructure only: nothing in the product reads this program or its output.
//
//	cd oracle/gnark_fixtures && go run . > ../../tests/golden/gnark_fixtures.json      (needs Go >= 1.24 and the module)
package main

import (
	"encoding/hex"
	"encoding/json"
	"math/big"
	"os"
	"unsafe"

	"github.com/consensys/gnark-crypto/ecc/bn254"
	"github.com/consensys/gnark-crypto/ecc/bn254/fp"
	"github.com/consensys/gnark-crypto/ecc/bn254/fr"
)

// SplitMix64, the generator of SURVEY.md 8d (oracle/bn254_ref.py SplitMix64): seed 0xB2000254 + config index.
type splitmix struct{ s uint64 }

func (g *splitmix) next() uint64 {
	g.s += 0x9E3779B97F4A7C15
	z := g.s
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9
	z = (z ^ (z >> 27)) * 0x94D049BB133111EB
	return z ^ (z >> 31)
}

// scalar = 4 consecutive draws as a 256-bit little-endian integer, mod r
func (g *splitmix) scalar() *big.Int {
	v := new(big.Int)
	for i := 0; i < 4; i++ {
		v.Or(v, new(big.Int).Lsh(new(big.Int).SetUint64(g.next()), uint(64*i)))
	}
	return v.Mod(v, fr.Modulus())
}

func raw[T any](p *T) string { return hex.EncodeToString(unsafe.Slice((*byte)(unsafe.Pointer(p)), unsafe.Sizeof(*p))) }
func le32(k *big.Int) string {
	var b [32]byte
	be := k.Bytes()
	for i, x := range be {
		b[len(be)-1-i] = x
	}
	return hex.EncodeToString(b[:])
}

type M = map[string]any

func main() {
	g := &splitmix{0xB2000254 + 100}
	_, _, g1, g2 := bn254.Generators()
	out := M{"gnark_crypto": "v0.19.0", "seed": "0xB2000254+100", "layout": "raw = in-memory struct bytes (4 x u64 LE limbs, Montgomery)"}

	edge := []*big.Int{big.NewInt(0), big.NewInt(1), big.NewInt(2), new(big.Int).Sub(fr.Modulus(), big.NewInt(1)),
		new(big.Int).Lsh(big.NewInt(1), 128)}
	var P []bn254.G1Affine
	var Q []bn254.G2Affine
	for i := 0; i < 6; i++ {
		var p bn254.G1Affine
		var q bn254.G2Affine
		p.ScalarMultiplicationBase(g.scalar())
		q.ScalarMultiplicationBase(g.scalar())
		P, Q = append(P, p), append(Q, q)
	}

	var pairs []M
	for i := range P {
		e, _ := bn254.Pair(P[i:i+1], Q[i:i+1])
		b := e.Bytes()
		pairs = append(pairs, M{"P": raw(&P[i]), "Q": raw(&Q[i]), "gt": raw(&e), "gt_bytes": hex.EncodeToString(b[:]), "gt_marshal": hex.EncodeToString(e.Marshal())})
	}
	out["pair"] = pairs
	e3, _ := bn254.Pair(P[:3], Q[:3])
	ml, _ := bn254.MillerLoop(P[:3], Q[:3])
	out["multi_pair"] = M{"k": 3, "P": raw(&P[0]) + raw(&P[1]) + raw(&P[2]), "Q": raw(&Q[0]) + raw(&Q[1]) + raw(&Q[2]), "gt": raw(&e3),
		"final_exp_of_miller_loop_equals_pair": bn254.FinalExponentiation(&ml) == e3}
	ok, _ := bn254.PairingCheck(P[:2], Q[:2])
	var negP bn254.G1Affine
	negP.Neg(&P[0])
	ok2, _ := bn254.PairingCheck([]bn254.G1Affine{P[0], negP}, []bn254.G2Affine{Q[0], Q[0]})
	out["pairing_check"] = []M{{"P": raw(&P[0]) + raw(&P[1]), "Q": raw(&Q[0]) + raw(&Q[1]), "ok": ok},
		{"P": raw(&P[0]) + raw(&negP), "Q": raw(&Q[0]) + raw(&Q[0]), "ok": ok2}}

	// FinalExponentiation of an arbitrary Fp12 element (not a Miller value)
	var z bn254.GT
	coeffs := (*[12]fp.Element)(unsafe.Pointer(&z)) // E12 = 12 fp.Element in memory order C0.B0.A0 ... C1.B2.A1
	for i := 0; i < 12; i++ {
		coeffs[i].SetBigInt(g.scalar()) // < r < p: a valid fp value
	}
	fe := bn254.FinalExponentiation(&z)
	out["final_exp"] = M{"in": raw(&z), "out": raw(&fe)}

	var g1m, g2m, gte []M
	scalars := append([]*big.Int{}, edge...)
	for i := 0; i < 4; i++ {
		scalars = append(scalars, g.scalar())
	}
	ePQ, _ := bn254.Pair(P[:1], Q[:1])
	for _, k := range scalars {
		var a bn254.G1Affine
		var b bn254.G2Affine
		var x bn254.GT
		a.ScalarMultiplication(&P[1], k)
		b.ScalarMultiplication(&Q[1], k)
		x.Exp(ePQ, k)
		g1m = append(g1m, M{"base": raw(&P[1]), "k": le32(k), "out": raw(&a)})
		g2m = append(g2m, M{"base": raw(&Q[1]), "k": le32(k), "out": raw(&b)})
		gte = append(gte, M{"x": raw(&ePQ), "k": le32(k), "out": raw(&x)})
	}
	var gb1 bn254.G1Affine
	var gb2 bn254.G2Affine
	gb1.ScalarMultiplicationBase(scalars[6])
	gb2.ScalarMultiplicationBase(scalars[6])
	out["g1_mul"], out["g2_mul"], out["gt_exp"] = g1m, g2m, gte
	out["mul_base"] = M{"k": le32(scalars[6]), "g1": raw(&gb1), "g2": raw(&gb2), "g1_gen": raw(&g1), "g2_gen": raw(&g2)}

	var s1, d1 bn254.G1Affine
	var s2, d2 bn254.G2Affine
	s1.Add(&P[2], &P[3])
	d1.Add(&P[2], &P[2])
	s2.Add(&Q[2], &Q[3])
	d2.Sub(&Q[2], &Q[2])
	out["add"] = M{"g1": M{"a": raw(&P[2]), "b": raw(&P[3]), "sum": raw(&s1), "dbl": raw(&d1)},
		"g2": M{"a": raw(&Q[2]), "b": raw(&Q[3]), "sum": raw(&s2), "a_minus_a": raw(&d2)}}
	var gm, gd, gi bn254.GT
	e1, _ := bn254.Pair(P[1:2], Q[1:2])
	gm.Mul(&ePQ, &e1)
	gd.Div(&ePQ, &e1)
	gi.Inverse(&ePQ)
	out["gt_ops"] = M{"a": raw(&ePQ), "b": raw(&e1), "mul": raw(&gm), "div": raw(&gd), "inv_a": raw(&gi)}

	// hash-to-curve with the reference's four domain separation tags (hash/hash_to.go:114,170,204,272)
	var hs []M
	for _, dst := range []string{"Hash String To Element In G1", "Hash Bytes To Element In G1", "Hash String To Element In G2", "Hash Bytes To Element In G2",
		"QUUX-V01-CS02-with-BN254G2_XMD:SHA-256_SVDW_RO_"} {
		for _, msg := range [][]byte{{}, []byte("abc"), []byte("batch-0"), make([]byte, 130)} {
			h1, _ := bn254.HashToG1(msg, []byte(dst))
			h2, _ := bn254.HashToG2(msg, []byte(dst))
			hs = append(hs, M{"dst": dst, "msg": hex.EncodeToString(msg), "g1": raw(&h1), "g2": raw(&h2)})
		}
	}
	out["hash_to_curve"] = hs

	// wire encodings
	var inf1 bn254.G1Affine
	var inf2 bn254.G2Affine
	enc := func(p *bn254.G1Affine, q *bn254.G2Affine) M {
		b1, b2 := p.Bytes(), q.Bytes()
		return M{"g1": raw(p), "g1_bytes": hex.EncodeToString(b1[:]), "g1_marshal": hex.EncodeToString(p.Marshal()),
			"g2": raw(q), "g2_bytes": hex.EncodeToString(b2[:]), "g2_marshal": hex.EncodeToString(q.Marshal())}
	}
	out["wire"] = []M{enc(&P[0], &Q[0]), enc(&P[4], &Q[4]), enc(&negP, &Q[5]), enc(&inf1, &inf2)}
	var frs []M
	for _, k := range scalars {
		var e fr.Element
		e.SetBigInt(k)
		b := e.Bytes()
		frs = append(frs, M{"value": le32(k), "raw": raw(&e), "bytes": hex.EncodeToString(b[:])})
	}
	out["fr"] = frs

	enc2 := json.NewEncoder(os.Stdout)
	enc2.SetIndent("", " ")
	if err := enc2.Encode(out); err != nil {
		panic(err)
	}
}
