"""the scalar-field feeders.  Reference behaviour restated with Python integers:
utils/compute_lagrange_basis.go:8-30, bibe/afp25_bibe/afp25_bibe_utils.go:14-43 (computePolynomialCoeffs) and the
quotient polynomial of bibe/afp25_bibe/afp25_bibe.go:369-383."""
import numpy as np
import pytest

from oracle import bn254_ref as o

R = o.R


def poly_from_roots(roots):
    """afp25_bibe_utils.go:14-43, literally: start from [1], multiply by (X - root) one root at a time."""
    c = [1]
    for r in roots:
        n = [0] * (len(c) + 1)
        for i, v in enumerate(c):
            n[i] = (n[i] - r * v) % R
            n[i + 1] = (n[i + 1] + v) % R
        c = n
    return c


def lagrange(i, s, x):
    """compute_lagrange_basis.go:8-30: factors with j == i (by VALUE) are skipped."""
    d = 1
    for j in s:
        if j != i:
            d = d * ((x - j) % R) % R * pow((i - j) % R, -1, R) % R
    return d


def test_fr_constants_and_host_helpers():
    """Host-side helpers of the library (no GPU): Montgomery conversion and the one-inversion Lagrange basis."""
    import __graft_entry__ as g

    g.build()
    from gopairingbasedcryptography_b200 import bn254

    rng = o.SplitMix64(4242)
    vals = [0, 1, 2, R - 1, R - 2, 1 << 128] + [rng.scalar() for _ in range(40)]
    e = bn254.fr_from_ints(vals)
    # gnark fr.Element: value * 2^256 mod r, little-endian limbs
    assert int.from_bytes(e[1].tobytes(), "little") == (1 << 256) % R
    assert [int.from_bytes(v.tobytes(), "little") for v in bn254.fr_to_scalars(e)] == vals
    assert bn254.fr_to_ints(e) == vals
    # BSW07's gate: children 1..100, x = 0 (access/tree/access_tree_node.go:151-158)
    for s, x in (([i + 1 for i in range(100)], 0), ([rng.scalar() for _ in range(17)], rng.scalar()), ([5], 9), ([3, 7, 3, 11], 0)):
        got = bn254.fr_to_ints(bn254.fr_lagrange_basis(bn254.fr_from_ints(s), bn254.fr_from_ints([x])))
        assert got == [lagrange(i, s, x) for i in s]
    # a 100-leaf gate's coefficients interpolate a degree-99 polynomial at 0
    coef = [rng.scalar() for _ in range(100)]
    ev = lambda t: sum(c * pow(t, k, R) for k, c in enumerate(coef)) % R
    s = list(range(1, 101))
    d = bn254.fr_to_ints(bn254.fr_lagrange_basis(bn254.fr_from_ints(s), bn254.fr_from_ints([0])))
    assert sum(di * ev(i) for di, i in zip(d, s)) % R == coef[0]


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 100, 1024])
def test_poly_from_roots_and_quotients_on_gpu(engine, n):
    from gopairingbasedcryptography_b200 import bn254

    rng = o.SplitMix64(9000 + n)
    roots = [10000 + 10 * i for i in range(n)] if n == 1024 else [rng.scalar() for _ in range(n)]  # afp25_bibe_test.go:381 id pattern
    f = engine.fr_poly_from_roots(bn254.fr_from_ints(roots))
    ref = poly_from_roots(roots)
    assert bn254.fr_to_ints(f) == ref
    pick = sorted(set([0, n - 1, n // 2] + [int(rng.next() % n) for _ in range(5)]))
    q = engine.fr_quotient_coeffs(f, bn254.fr_from_ints([roots[i] for i in pick]))
    assert q.shape == (len(pick), n, 32)
    for row, i in zip(q, pick):
        # the reference expands the polynomial of the remaining roots (afp25_bibe.go:372-383)
        want = poly_from_roots(roots[:i] + roots[i + 1:])
        assert [int.from_bytes(v.tobytes(), "little") for v in row] == want
