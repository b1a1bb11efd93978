// BLS sign + verify of 1024 random messages (BASELINE configs[0]) on the batch entry points: the reference's
// signature/bls01_signature flow with the three loops replaced by three calls.
package main

import (
	"crypto/rand"
	"fmt"
	"math/big"

	"github.com/mmsyan/GoPairingBasedCryptography/b200/bn254"
	"github.com/mmsyan/GoPairingBasedCryptography/b200/bn254/fr"
)

func main() {
	const n = 1024
	dst := []byte("Hash Bytes To Element In G2") // hash/hash_to.go:272
	msgs := make([][]byte, n)
	for i := range msgs {
		msgs[i] = make([]byte, 32)
		rand.Read(msgs[i])
	}
	var x fr.Element
	x.SetRandom()
	sk := x.BigInt(new(big.Int))
	var pk bn254.G1Affine
	pk.ScalarMultiplicationBase(sk) // bls_signature.go:45

	hm, err := bn254.HashToG2Batch(msgs, dst) // bls_signature.go:60,73
	if err != nil {
		panic(err)
	}
	sks := make([]bn254.Scalar, n)
	for i := range sks {
		sks[i] = bn254.ScalarFromBig(sk)
	}
	sig, err := bn254.ScalarMulBatchG2(hm, sks) // bls_signature.go:63
	if err != nil {
		panic(err)
	}
	ok, err := bn254.BLSVerifyBatch(&pk, hm, sig) // bls_signature.go:71-89
	if err != nil {
		panic(err)
	}
	good := 0
	for _, b := range ok {
		if b {
			good++
		}
	}
	fmt.Printf("%d / %d signatures verify on %d GPU(s)\n", good, n, bn254.DeviceCount())
}
