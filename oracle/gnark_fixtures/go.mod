// Fixture generator: runs the UNMODIFIED gnark-crypto the reference pins (/root/reference/go.mod:5,
// go.sum:3-4 h1:zXCqeY2txSaMl6G5wFpZzMWJU9HPNh8qxPnYJ1BL9vA=) on seeded inputs.
module gnarkfixtures

go 1.24

require github.com/consensys/gnark-crypto v0.19.0
