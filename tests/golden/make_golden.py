"""Generate tests/golden/bn254_vectors.json from the DEFINITIONAL Python oracle (oracle/bn254_ref.py).

Run from the repo root:  python tests/golden/make_golden.py
The reference holds no golden vectors for this path and gnark cannot run in this image (SURVEY.md
§8c), so these are "oracle-generated" vectors: they freeze today's definitional answers so that the
C restatement, the host-emulated device code and the CUDA path are all checked against the same
committed bytes.  Inputs are seeded (SplitMix64, 0xB2000254).
"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from oracle import bn254_ref as o  # noqa: E402


def main():
    rng = o.SplitMix64(0xB2000254)
    v = {"comment": "oracle-generated (definitional big-int); gnark memory layout, hex", "pair": [], "multi_pair": [],
         "g1_mul": [], "g2_mul": [], "gt_exp": [], "g1_add": [], "g2_add": []}
    e_gen = o.pair([o.G1_GEN], [o.G2_GEN])
    v["e_g1_g2"] = o.gt_to_bytes(e_gen).hex()
    v["e_g1_g2_literal_final_exp"] = o.gt_to_bytes(o.final_exponentiation_literal(o.miller_loop([o.G1_GEN], [o.G2_GEN]))).hex()
    pts = []
    for _ in range(6):
        a, b = rng.scalar(), rng.scalar()
        P, Q = o.g1_mul(o.G1_GEN, a), o.g2_mul(o.G2_GEN, b)
        pts.append((P, Q))
        v["pair"].append({"P": o.g1_to_bytes(P).hex(), "Q": o.g2_to_bytes(Q).hex(), "gt": o.gt_to_bytes(o.pair([P], [Q])).hex()})
    # infinity members are skipped
    v["pair"].append({"P": o.g1_to_bytes(None).hex(), "Q": o.g2_to_bytes(pts[0][1]).hex(), "gt": o.gt_to_bytes(o.FP12_ONE).hex()})
    v["pair"].append({"P": o.g1_to_bytes(pts[0][0]).hex(), "Q": o.g2_to_bytes(None).hex(), "gt": o.gt_to_bytes(o.FP12_ONE).hex()})
    for k in (2, 3, 5):
        Ps, Qs = [p for p, _ in pts[:k]], [q for _, q in pts[:k]]
        v["multi_pair"].append({"k": k, "P": b"".join(o.g1_to_bytes(p) for p in Ps).hex(),
                                "Q": b"".join(o.g2_to_bytes(q) for q in Qs).hex(), "gt": o.gt_to_bytes(o.pair(Ps, Qs)).hex()})
    # e(P,Q) e(-P,Q) = 1  (BLS-verify shape)
    P, Q = pts[1]
    v["multi_pair"].append({"k": 2, "P": (o.g1_to_bytes(P) + o.g1_to_bytes(o.g1_neg(P))).hex(),
                            "Q": (o.g2_to_bytes(Q) * 2).hex(), "gt": o.gt_to_bytes(o.FP12_ONE).hex()})
    base1, base2 = pts[2]
    for k in o.EDGE_SCALARS + [rng.scalar() for _ in range(3)] + [(1 << 256) - 1]:
        v["g1_mul"].append({"base": o.g1_to_bytes(base1).hex(), "k": o.scalar_to_bytes(k).hex(), "out": o.g1_to_bytes(o.g1_mul(base1, k)).hex()})
        v["g2_mul"].append({"base": o.g2_to_bytes(base2).hex(), "k": o.scalar_to_bytes(k).hex(), "out": o.g2_to_bytes(o.g2_mul(base2, k)).hex()})
    x = o.pair([pts[3][0]], [pts[3][1]])
    nonsub = o.fp12_from_coeffs([rng.fp() for _ in range(12)])  # GT.SetRandom-style element outside the subgroup
    for base in (x, nonsub):
        for k in [0, 1, 2, o.R - 1, rng.scalar(), (1 << 256) - 1]:
            v["gt_exp"].append({"x": o.gt_to_bytes(base).hex(), "k": o.scalar_to_bytes(k).hex(), "out": o.gt_to_bytes(o.gt_exp(base, k)).hex()})
    v["gt_mul"] = {"a": o.gt_to_bytes(x).hex(), "b": o.gt_to_bytes(nonsub).hex(), "mul": o.gt_to_bytes(o.fp12_mul(x, nonsub)).hex(),
                   "div": o.gt_to_bytes(o.gt_div(x, nonsub)).hex()}
    A, B = pts[4][0], pts[5][0]
    for a, b in ((A, B), (A, A), (A, o.g1_neg(A)), (None, B), (A, None), (None, None)):
        v["g1_add"].append({"a": o.g1_to_bytes(a).hex(), "b": o.g1_to_bytes(b).hex(), "out": o.g1_to_bytes(o.g1_add(a, b)).hex()})
    A, B = pts[4][1], pts[5][1]
    for a, b in ((A, B), (A, A), (A, o.g2_neg(A)), (None, B), (A, None), (None, None)):
        v["g2_add"].append({"a": o.g2_to_bytes(a).hex(), "b": o.g2_to_bytes(b).hex(), "out": o.g2_to_bytes(o.g2_add(a, b)).hex()})
    f = o.miller_loop([pts[0][0]], [pts[0][1]])
    v["final_exp"] = {"in": o.gt_to_bytes(f).hex(), "out": o.gt_to_bytes(o.final_exponentiation_literal(f)).hex()}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bn254_vectors.json")
    with open(path, "w") as fh:
        json.dump(v, fh, indent=1)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
