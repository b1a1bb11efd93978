// Drop-in replacement for the gnark-crypto `ecc/bn254` call surface the schemes of
// mmsyan/GoPairingBasedCryptography use, backed by libbn254_b200.so (include/bn254_b200.h).
// gnark-crypto stays a dependency: fr.Element is gnark's own type (host-side scalar arithmetic is out of the
// engine's scope), and byte encodings / String() delegate to gnark on identical memory layouts.
module github.com/mmsyan/GoPairingBasedCryptography/b200

go 1.24

require github.com/consensys/gnark-crypto v0.19.0
