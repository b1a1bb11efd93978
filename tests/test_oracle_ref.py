"""Definitional Python oracle: algebraic properties that pin the pairing without gnark
(SURVEY.md §8c "how the oracle earns trust")."""
from oracle import bn254_ref as o


def test_curve_constants():
    assert o.g1_on_curve(o.G1_GEN) and o.g2_on_curve(o.G2_GEN)
    assert o.g1_mul(o.G1_GEN, o.R - 1) == o.g1_neg(o.G1_GEN)
    assert o.g2_add(o.g2_mul(o.G2_GEN, o.R - 1), o.G2_GEN) is None
    assert (o.LAMBDA_GLV**2 + o.LAMBDA_GLV + 1) % o.R == 0
    assert o.FE_COFACTOR.bit_length() == 190 and o.FE_EXPONENT.bit_length() == 2980


def test_frobenius_is_p_power():
    f = o.miller_loop([o.G1_GEN], [o.G2_GEN])
    assert o.fp12_frobenius(f, 1) == o.fp12_pow(f, o.P)
    assert o.fp12_frobenius(f, 2) == o.fp12_pow(f, o.P**2)
    assert o.fp12_frobenius(f, 3) == o.fp12_pow(f, o.P**3)
    assert o.g2_frobenius(o.G2_GEN, 1) == o.g2_mul(o.G2_GEN, o.P % o.R)


def test_final_exp_literal_equals_structured():
    f = o.miller_loop([o.g1_mul(o.G1_GEN, 5)], [o.g2_mul(o.G2_GEN, 7)])
    assert o.final_exponentiation(f) == o.final_exponentiation_literal(f)


def test_bilinearity_nondegeneracy_order():
    e = o.pair([o.G1_GEN], [o.G2_GEN])
    assert e != o.FP12_ONE
    assert o.fp12_pow(e, o.R) == o.FP12_ONE
    a, b = 0x1234567890ABCDEF1234567890ABCDEF, 0xFEDCBA0987654321FEDCBA0987654321
    assert o.pair([o.g1_mul(o.G1_GEN, a)], [o.g2_mul(o.G2_GEN, b)]) == o.fp12_pow(e, a * b % o.R)
    # additivity in the second argument and product form
    Q1, Q2 = o.g2_mul(o.G2_GEN, 11), o.g2_mul(o.G2_GEN, 31)
    lhs = o.pair([o.G1_GEN], [o.g2_add(Q1, Q2)])
    assert lhs == o.fp12_mul(o.pair([o.G1_GEN], [Q1]), o.pair([o.G1_GEN], [Q2]))
    assert lhs == o.pair([o.G1_GEN, o.G1_GEN], [Q1, Q2])


def test_final_exp_invariant_under_line_rescaling():
    # multiplying the Miller value by a proper-subfield element must not change the pairing
    f = o.miller_loop([o.G1_GEN], [o.G2_GEN])
    sub = ((((3, 5), (7, 11), (13, 17))), o.FP6_ZERO)  # an Fp6 element embedded in Fp12
    assert o.final_exponentiation(o.fp12_mul(f, sub)) == o.final_exponentiation(f)


def test_infinity_and_errors():
    import pytest

    assert o.pair([None], [o.G2_GEN]) == o.FP12_ONE
    assert o.pair([o.G1_GEN], [None]) == o.FP12_ONE
    with pytest.raises(ValueError):
        o.pair([], [])
    with pytest.raises(ValueError):
        o.pair([o.G1_GEN], [])
    P = o.g1_mul(o.G1_GEN, 9)
    assert o.pairing_check([P, o.g1_neg(P)], [o.G2_GEN, o.G2_GEN])
    assert not o.pairing_check([P, P], [o.G2_GEN, o.G2_GEN])


def test_gt_exp_semantics():
    e = o.pair([o.G1_GEN], [o.G2_GEN])
    assert o.gt_exp(e, 0) == o.FP12_ONE
    assert o.fp12_mul(o.gt_exp(e, -5), o.gt_exp(e, 5)) == o.FP12_ONE


def test_fr_polynomial_kat():
    """The only known-answer test in the reference: (x-1)(x-2) -> [2, -3, 1] over Fr
    (bibe/gwww25_bibe/gwww25_bibe_test.go:400-428)."""
    coeffs = [1]
    for root in (1, 2):
        nxt = [0] * (len(coeffs) + 1)
        for i, c in enumerate(coeffs):
            nxt[i + 1] = (nxt[i + 1] + c) % o.R
            nxt[i] = (nxt[i] - c * root) % o.R
        coeffs = nxt
    assert coeffs == [2, o.R - 3, 1]


def test_montgomery_layout_round_trip():
    assert o.fp_from_mont_bytes(o.fp_to_mont_bytes(12345)) == 12345
    assert o.fp_to_mont_bytes(1).hex().startswith("9d0d8fc5")  # R mod p, low limb 0x...c58f0d9d little-endian
    e = o.pair([o.G1_GEN], [o.G2_GEN])
    assert o.gt_from_bytes(o.gt_to_bytes(e)) == e
    assert o.g2_from_bytes(o.g2_to_bytes(o.G2_GEN)) == o.G2_GEN
