"""B200-native batched BN254 pairing engine (sm_100a CUDA behind a C ABI).

Public surface: :mod:`gopairingbasedcryptography_b200.bn254` mirrors the gnark-crypto ``bn254``
names the reference's schemes call (``Pair``, ``PairingCheck``, ``MillerLoop``,
``FinalExponentiation``, ``G1Affine/G2Affine.ScalarMultiplication``, ``GT.Exp`` ...) plus the batch
entry points that reach the GPU.  There is no CPU fallback.
"""
from . import bn254  # noqa: F401

__all__ = ["bn254"]
