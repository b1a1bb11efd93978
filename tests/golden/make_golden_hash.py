"""Golden vectors for hash-to-curve, generated from the definitional oracle (oracle/hash_to_curve_ref.py):
   python tests/golden/make_golden_hash.py  ->  tests/golden/hash_to_curve_vectors.json
The expand_message_xmd entries are the RFC 9380 appendix K.1 vectors (published values, not generated)."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import bn254_ref as o  # noqa: E402
from oracle import hash_to_curve_ref as h  # noqa: E402

msgs = [b"", b"abc", b"abcdef0123456789", b"q128_" + b"q" * 128, b"a512_" + b"a" * 512, bytes(range(256))]
out = {
    "rfc9380_k1_expand_message_xmd_sha256": {
        "dst": "QUUX-V01-CS02-with-expander-SHA256-128",
        "vectors": [
            {"msg": "", "len": 32, "uniform_bytes": "68a985b87eb6b46952128911f2a4412bbc302a9d759667f87f7a21d803f07235"},
            {"msg": "abc", "len": 32, "uniform_bytes": "d8ccab23b5985ccea865c6c97b6e5b8350e794e603b4b97902f53a8a0d605615"},
        ],
    },
    "g1": [], "g2": [],
}
for dst in (h.DST_BYTES_G1, h.DST_STRING_G1):
    for m in msgs:
        out["g1"].append({"msg": m.hex(), "dst": dst.decode(), "point": o.g1_to_bytes(h.hash_to_g1(m, dst)).hex()})
for dst in (h.DST_BYTES_G2, h.DST_STRING_G2):
    for m in msgs:
        out["g2"].append({"msg": m.hex(), "dst": dst.decode(), "point": o.g2_to_bytes(h.hash_to_g2(m, dst)).hex()})
with open(os.path.join(HERE, "hash_to_curve_vectors.json"), "w") as f:
    json.dump(out, f, indent=1)
print("wrote", len(out["g1"]), "G1 and", len(out["g2"]), "G2 vectors")
