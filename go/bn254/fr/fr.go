// Package fr re-exports gnark's scalar field: host-side Fr arithmetic is outside the engine's scope (SURVEY.md 8a
// row 11 -- the schemes use fr.Element as map keys and for a few dozen operations per call), so the type IS
// gnark's (an alias keeps every method, ==, and map-key use).  What this package adds are the batch feeders that
// replace the reference's quadratic / inversion-heavy helpers.
package fr

/*
#cgo CFLAGS: -I${SRCDIR}/../../../include
#cgo LDFLAGS: -L${SRCDIR}/../../../gopairingbasedcryptography_b200/lib -lbn254_b200
#include "bn254_b200.h"
*/
import "C"

import (
	"unsafe"

	gfr "github.com/consensys/gnark-crypto/ecc/bn254/fr"
)

type Element = gfr.Element

const (
	Limbs = gfr.Limbs
	Bits  = gfr.Bits
	Bytes = gfr.Bytes
)

var (
	NewElement = gfr.NewElement
	Modulus    = gfr.Modulus
	One        = gfr.One
)

// LagrangeBasis returns Delta_{s[i],S}(x) for every i with ONE field inversion (Montgomery's trick); the reference's
// utils.ComputeLagrangeBasis (utils/compute_lagrange_basis.go:8-30) inverts once per factor -- 9 900 inversions for a
// 100-leaf gate (access/tree/access_tree_node.go:151-158).  Factors with s[j] == s[i] BY VALUE are skipped, exactly
// as the reference does.
func LagrangeBasis(s []Element, x Element) []Element {
	out := make([]Element, len(s))
	if len(s) > 0 {
		C.bn254_fr_lagrange_basis(unsafe.Pointer(&s[0]), C.size_t(len(s)), unsafe.Pointer(&x), unsafe.Pointer(&out[0]))
	}
	return out
}

// ComputeLagrangeBasis keeps the reference's signature for call sites that want one coefficient.
func ComputeLagrangeBasis(i Element, s []Element, x Element) Element {
	for k, b := range LagrangeBasis(s, x) {
		if s[k] == i {
			return b
		}
	}
	// i not in S: the reference's product runs over all of S
	num, den := One(), One()
	for _, j := range s {
		var n, d Element
		n.Sub(&x, &j)
		d.Sub(&i, &j)
		num.Mul(&num, &n)
		den.Mul(&den, &d)
	}
	den.Inverse(&den)
	num.Mul(&num, &den)
	return num
}
