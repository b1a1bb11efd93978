"""CPU-side checks of the drop-in boundary: the shared library loads and exports every symbol that
include/bn254_b200.h declares, and the host mirror's argument handling follows gnark's conventions.
No compute calls are made (there is no GPU here and no CPU fallback in the library)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "bn254_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bn254_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g

    g.build()
    from gopairingbasedcryptography_b200 import _native

    lib = _native.lib()
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), s
    assert sorted(_native.SYMBOLS) == syms


def test_no_device_means_loud_failure():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from gopairingbasedcryptography_b200 import bn254

    with pytest.raises(bn254.EngineError):
        bn254.Engine(0)


def test_host_mirror_conventions():
    from gopairingbasedcryptography_b200 import bn254

    assert bn254.GT().IsZero()  # zero value of GT is 0, not 1
    assert bn254.G1Affine().IsInfinity() and bn254.G2Affine().IsInfinity()
    with pytest.raises(ValueError, match="invalid inputs sizes"):
        bn254._pack([], [])
    with pytest.raises(ValueError, match="invalid inputs sizes"):
        bn254._pack([bn254.G1Affine()], [])
    assert bn254._norm_scalar(-3) == (True, 3)
    assert bn254._norm_scalar(bn254.R_MOD + 5) == (False, 5)
    # Neg on raw Montgomery bytes: y -> p - y, infinity stays infinity
    g = bn254.Generators()[2]
    n = bn254.G1Affine().Neg(g)
    y, ny = int.from_bytes(g.raw[32:], "little"), int.from_bytes(n.raw[32:], "little")
    assert (y + ny) % bn254.P_MOD == 0 and n.raw[:32] == g.raw[:32]
    assert bn254.G1Affine().Neg(bn254.G1Affine()).IsInfinity()
    assert bn254.scalars_to_bytes([1, 2**255]).shape == (2, 32)
    with pytest.raises(ValueError):
        bn254.scalars_to_bytes([2**256])


def test_generators_match_oracle():
    from gopairingbasedcryptography_b200 import bn254
    from oracle import bn254_ref as o

    _, _, g1, g2 = bn254.Generators()
    assert g1.raw == o.g1_to_bytes(o.G1_GEN) and g2.raw == o.g2_to_bytes(o.G2_GEN)


def test_device_constants_match_oracle():
    """gen_constants.py derives its tables independently of oracle/; cross-check them."""
    from oracle import bn254_ref as o

    txt = open(os.path.join(ROOT, "gopairingbasedcryptography_b200", "csrc", "bn254_constants.cuh")).read()

    def limbs(s):
        return sum(int(x, 16) << (32 * i) for i, x in enumerate(re.findall(r"0x([0-9a-f]{8})u", s)))

    m = re.search(r"GAMMA1\[6\] = \{(.*?)\};", txt, re.S).group(1)
    rows = [r for r in m.strip().split("\n") if r.strip()]
    for j, row in enumerate(rows):
        a0, a1 = re.findall(r"\{\{([^{}]*)\}\}", row)
        assert limbs(a0) == o.GAMMA1[j][0] * o.MONT_R % o.P and limbs(a1) == o.GAMMA1[j][1] * o.MONT_R % o.P
    m = re.search(r"TWIST_3B = \{(.*?)\};", txt).group(1)
    a0, a1 = re.findall(r"\{\{([^{}]*)\}\}", m)
    b3 = o.fp2_scale(o.B2, 3)
    assert limbs(a0) == b3[0] * o.MONT_R % o.P and limbs(a1) == b3[1] * o.MONT_R % o.P
    beta = limbs(re.search(r"GLV_BETA = \{\{(.*?)\}\}", txt).group(1)) * pow(o.MONT_R, -1, o.P) % o.P
    assert pow(beta, 3, o.P) == 1 and beta != 1
