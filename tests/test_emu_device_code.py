"""The DEVICE algorithms (csrc/*.cuh) compiled for the host with an emulated PTX carry flag, checked
bit-for-bit against the oracle.  This is a test harness for the CUDA sources on a box without a GPU;
it is not a product code path (the shipped .so has no CPU fallback)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import bn254_ref as o
from oracle import port

import common

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emu", "emu_bn254.cpp")
# BN254_EMU_FLAGS: extra -D flags (e.g. "-DBN254_STAGED" / "-DBN254_LEGACY") to check a non-default device build
EMU_FLAGS = os.environ.get("BN254_EMU_FLAGS", "").split()
SO = os.path.join(HERE, "emu", "_emu_bn254%s.so" % ("_" + "_".join(f.lstrip("-D") for f in EMU_FLAGS) if EMU_FLAGS else ""))
CSRC = os.path.join(HERE, "..", "gopairingbasedcryptography_b200", "csrc")


@pytest.fixture(scope="module")
def emu():
    from gopairingbasedcryptography_b200 import _build

    _build.ensure_generated()  # the warp-VM programs are generated sources
    deps = [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".inc"))]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-DBN254_OOL_ADDS", "-DBN254_OOL_FPMUL"] + EMU_FLAGS +
                              ["-o", SO, SRC])
    return ctypes.CDLL(SO)


def vp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


sz = ctypes.c_size_t


def test_fp_arithmetic(emu):
    rng = o.SplitMix64(11)
    edge = [0, 1, o.P - 1, o.P - 2, 1 << 253, (1 << 32) - 1, 1 << 224]
    n = 512
    A = edge + [rng.fp() for _ in range(n - len(edge))]
    B = [rng.fp() for _ in range(n - len(edge))] + edge
    enc = lambda v: np.frombuffer(b"".join(x.to_bytes(32, "little") for x in v), dtype=np.uint8).copy()
    dec = lambda a: [int.from_bytes(a[32 * i:32 * i + 32].tobytes(), "little") for i in range(len(a) // 32)]
    a, b, z = enc(A), enc(B), np.zeros(32 * n, np.uint8)
    rinv = pow(1 << 256, -1, o.P)
    emu.emu_fp_mul(vp(a), vp(b), sz(n), vp(z))
    assert dec(z) == [x * y * rinv % o.P for x, y in zip(A, B)]
    emu.emu_fp_add(vp(a), vp(b), sz(n), vp(z))
    assert dec(z) == [(x + y) % o.P for x, y in zip(A, B)]
    emu.emu_fp_sub(vp(a), vp(b), sz(n), vp(z))
    assert dec(z) == [(x - y) % o.P for x, y in zip(A, B)]
    emu.emu_fp_half(vp(a), sz(n), vp(z))
    assert dec(z) == [x * pow(2, -1, o.P) % o.P for x in A]
    emu.emu_fp_inv(vp(a), sz(16), vp(z))
    assert dec(z)[:16] == [(pow(x * rinv, -1, o.P) * (1 << 256) % o.P if x else 0) for x in A[:16]]
    # inputs up to 2p are tolerated by the Montgomery product (used by lazy-reduction paths)
    A2 = [rng.u256() % (2 * o.P) for _ in range(n)]
    B2 = [rng.u256() % (2 * o.P) for _ in range(n)]
    a, b = enc(A2), enc(B2)
    emu.emu_fp_mul(vp(a), vp(b), sz(n), vp(z))
    assert dec(z) == [x * y * rinv % o.P for x, y in zip(A2, B2)]


def test_pairing_paths(emu):
    n = 6
    P, Q, _, _ = common.points(n)
    ref = port.pair_batch(P, Q, n)
    out = np.zeros(384 * n, np.uint8)
    emu.emu_multi_pair(vp(P), vp(Q), sz(n), sz(1), 1, vp(out))
    assert (out == ref).all()
    for k in (2, 3, 6):
        m = n // k
        out = np.zeros(384 * m, np.uint8)
        emu.emu_multi_pair(vp(P), vp(Q), sz(m), sz(k), 1, vp(out))
        assert (out == port.multi_pair_batch(P, Q, m, k)).all()
    ml = np.zeros(384 * n, np.uint8)
    emu.emu_multi_pair(vp(P), vp(Q), sz(n), sz(1), 0, vp(ml))
    fe = np.zeros(384 * n, np.uint8)
    emu.emu_final_exp(vp(ml), sz(n), vp(fe))
    assert (fe == ref).all()
    assert (fe == port.final_exp_batch(ml, n)).all()


def test_pairing_infinity_and_check(emu):
    n = 4
    P, Q, _, _ = common.points(n, seed=21)
    P2, Q2 = common.with_infinities(P, Q)
    out = np.zeros(384 * n, np.uint8)
    emu.emu_multi_pair(vp(P2), vp(Q2), sz(n), sz(1), 1, vp(out))
    assert (out == port.pair_batch(P2, Q2, n)).all()
    one = o.gt_to_bytes(o.FP12_ONE)
    assert out[:384].tobytes() == one and out[768:1152].tobytes() == one
    # check: e(P,Q) e(-P,Q) == 1
    Pn = np.frombuffer(o.g1_to_bytes(o.g1_neg(o.g1_from_bytes(P[:64].tobytes()))), dtype=np.uint8)
    PP = np.concatenate([P[:64], Pn, P[:64], P[:64]])
    QQ = np.concatenate([Q[:128]] * 4)
    ok = np.zeros(2, np.uint8)
    emu.emu_multi_pair(vp(PP), vp(QQ), sz(2), sz(2), 2, vp(ok))
    assert list(ok) == [1, 0]


def test_groups_and_gt(emu):
    n = 10
    ks = common.scalars(n)
    sb = common.scalar_bytes(ks)
    P, Q, _, _ = common.points(n, seed=31)
    out = np.zeros(64 * n, np.uint8)
    emu.emu_g1_mul(vp(P), sz(1), vp(sb), sz(n), vp(out))
    assert (out == port.g1_mul_batch(P, sb, n)).all()
    out = np.zeros(128 * n, np.uint8)
    emu.emu_g2_mul(vp(Q), sz(1), vp(sb), sz(n), vp(out))
    assert (out == port.g2_mul_batch(Q, sb, n)).all()
    Pr, Qr = np.roll(P, 64), np.roll(Q, 128)
    out = np.zeros(64 * n, np.uint8)
    emu.emu_g1_add(vp(P), vp(Pr), sz(n), vp(out))
    assert (out == port.g1_add_batch(P, Pr, n)).all()
    emu.emu_g1_add(vp(P), vp(P), sz(n), vp(out))
    assert (out == port.g1_add_batch(P, P, n)).all()
    out = np.zeros(128 * n, np.uint8)
    emu.emu_g2_add(vp(Q), vp(Qr), sz(n), vp(out))
    assert (out == port.g2_add_batch(Q, Qr, n)).all()
    m = 4
    gt = port.pair_batch(P[:64 * m], Q[:128 * m], m)
    out = np.zeros(384 * m, np.uint8)
    emu.emu_gt_exp(vp(gt), sz(1), vp(sb), sz(m), vp(out))
    assert (out == port.gt_exp_batch(gt, sb[:32 * m], m)).all()
    gt2 = np.roll(gt, 384)
    emu.emu_gt_mul(vp(gt), vp(gt2), sz(m), 0, vp(out))
    assert (out == port.gt_mul_batch(gt, gt2, m)).all()
    emu.emu_gt_mul(vp(gt), vp(gt2), sz(m), 1, vp(out))
    assert (out == port.gt_div_batch(gt, gt2, m)).all()
    emu.emu_gt_sqr(vp(gt), sz(m), 1, vp(out))
    assert (out == port.gt_sqr_batch(gt, m)).all()


def test_tower_vm_programs_on_host(emu):
    """The lane-group kernels' generated micro-op programs (vmgen.py), run by the C++ interpreter (vm.cuh)
    with lock-step round semantics: pair, miller-only and final-exp-only programs (K = 3, the in-tree lane width)."""
    n = 3
    P, Q, _, _ = common.points(n, seed=123)
    ref = port.pair_batch(P, Q, n)
    for mode in (0,):
        out = np.zeros(384 * n, np.uint8)
        emu.emu_vm(vp(P), vp(Q), sz(n), mode, vp(out))
        assert (out == ref).all()
    ml = np.zeros(384 * n, np.uint8)
    emu.emu_vm(vp(P), vp(Q), sz(n), 1, vp(ml))
    fe = np.zeros(384 * n, np.uint8)
    emu.emu_vm(vp(ml), None, sz(n), 2, vp(fe))
    assert (fe == ref).all()
    # final exponentiation of an arbitrary (non-Miller) Fp12 input
    rng = o.SplitMix64(77)
    x = np.frombuffer(b"".join(o.fp_to_mont_bytes(rng.fp()) for _ in range(12)), dtype=np.uint8).copy()
    emu.emu_vm(vp(x), None, sz(1), 2, vp(fe))
    assert (fe[:384] == port.final_exp_batch(x, 1)).all()


def test_vmgen_python_evaluator_matches_oracle():
    """The generator's own integer evaluator on the scheduled K=3 and K=6 programs (independent of C++)."""
    import sys

    sys.path.insert(0, CSRC)
    import vmgen as g

    rng = o.SplitMix64(5)
    Pt, Qt = o.g1_mul(o.G1_GEN, rng.scalar()), o.g2_mul(o.G2_GEN, rng.scalar())
    for K in (3, 6):
        words, meta = g.build_pair_program(K, window=g.WINDOW[K], cold_lifetime=g.COLD_LIFETIME)
        ins = meta["in_slots"]
        slots = g.evaluate(words, K, 256, {ins[0]: Pt, ins[1]: Qt[0], ins[2]: Qt[1]})
        out = [slots[s] for s in meta["out_slots"]]
        assert ((out[0], out[1], out[2]), (out[3], out[4], out[5])) == o.pair([Pt], [Qt])
        assert meta["nslots"] <= 48 and meta["ncold"] <= 64


def test_glv_and_fixed_base_scalar_mul(emu):
    """2-dimensional GLV (any 256-bit scalar, incl. >= r) and the 32x255 fixed-base window path."""
    ks = common.scalars(24) + [(1 << 256) - 1, o.R, o.R + 1, o.LAMBDA_GLV - 1, o.LAMBDA_GLV + 1, 1 << 255]
    n = len(ks)
    sb = common.scalar_bytes(ks)
    P, Q, _, _ = common.points(n, seed=31)
    P[64 * 9:64 * 10] = 0
    out = np.zeros(64 * n, np.uint8)
    emu.emu_g1_mul_glv(vp(P), sz(1), vp(sb), sz(n), vp(out))
    assert (out == port.g1_mul_batch(P, sb, n)).all()
    out = np.zeros(128 * n, np.uint8)
    emu.emu_g2_mul_glv(vp(Q), sz(1), vp(sb), sz(n), vp(out))
    assert (out == port.g2_mul_batch(Q, sb, n)).all()
    # G2 by the 4-dimensional GLS ladder (psi endomorphism, 67-bit sub-scalars, 15-entry table), incl. an infinity base
    Q2 = Q.copy()
    Q2[128 * 5:128 * 6] = 0
    lam4 = o.P % o.R
    ks4 = ks + [lam4, lam4 - 1, lam4 * lam4 % o.R, pow(lam4, 3, o.R), (1 << 67) - 1, 1 << 66]
    sb4 = common.scalar_bytes(ks4)
    Q4 = np.concatenate([Q2, Q2[: 128 * (len(ks4) - n)]])
    out = np.zeros(128 * len(ks4), np.uint8)
    emu.emu_g2_mul_gls4(vp(Q4), sz(1), vp(sb4), sz(len(ks4)), vp(out))
    assert (out == port.g2_mul_batch(Q4, sb4, len(ks4))).all()
    g1, _ = port.generators()
    out = np.zeros(64 * n, np.uint8)
    emu.emu_g1_mul_fixed(vp(g1), vp(sb), sz(n), vp(out))
    assert (out == port.g1_mul_base_batch(g1, sb, n)).all()


def test_lazy_fp2_product(emu):
    rng = o.SplitMix64(3)
    n = 400
    A = [rng.fp() for _ in range(2 * n)]
    B = [rng.fp() for _ in range(2 * n)]
    for i in range(8):
        A[i] = o.P - 1
        B[i] = o.P - 1
    A[8:12] = [0, 1, o.P - 1, 0]
    enc = lambda v: np.frombuffer(b"".join(o.fp_to_mont_bytes(x) for x in v), dtype=np.uint8).copy()
    a, b, z = enc(A), enc(B), np.zeros(64 * n, np.uint8)
    emu.emu_fp2_mul_lazy(vp(a), vp(b), sz(n), vp(z))
    for i in range(n):
        r = o.fp2_mul((A[2 * i], A[2 * i + 1]), (B[2 * i], B[2 * i + 1]))
        assert (o.fp_from_mont_bytes(z[64 * i:64 * i + 32].tobytes()), o.fp_from_mont_bytes(z[64 * i + 32:64 * i + 64].tobytes())) == r


def test_gt_exp_fixed_windows(emu):
    """Lane-uniform fixed-window ladders: generic (2-bit) and cyclotomic (signed 3-bit) vs the oracle's ladder."""
    ks = common.scalars(12) + [(1 << 256) - 1, 1 << 255, (1 << 255) + 1, 5, 7, 3, 4, 8, (1 << 256) - 8,
                               int("7" * 64, 16), int("4" * 64, 16), int("c" * 64, 16), int("b6d" * 21 + "b", 16)]
    n = len(ks)
    sb = common.scalar_bytes(ks)
    P, Q, _, _ = common.points(n, seed=55)
    gt = port.pair_batch(P, Q, n)
    ref = port.gt_exp_batch(gt, sb, n)
    out = np.zeros(384 * n, np.uint8)
    emu.emu_gt_cyclo_exp(vp(gt), sz(1), vp(sb), sz(n), vp(out))
    assert (out == ref).all()
    emu.emu_gt_exp(vp(gt), sz(1), vp(sb), sz(n), vp(out))
    assert (out == ref).all()
    # generic ladder on an element outside the cyclotomic subgroup
    rng = o.SplitMix64(99)
    x = np.frombuffer(b"".join(o.fp_to_mont_bytes(rng.fp()) for _ in range(12)), dtype=np.uint8).copy()
    emu.emu_gt_exp(vp(x), sz(0), vp(sb), sz(4), vp(out))
    assert (out[:384 * 4] == port.gt_exp_base_batch(x, sb[:128], 4)).all()


def test_precomputed_g2_lines(emu):
    """Line tables of fixed G2 points + line-based Miller product == the ordinary multi-pairing (incl. infinity)."""
    n, m = 3, 11             # two chunks of table points (8 + 3): lines are applied two at a time, left-overs alone
    _, Q, _, _ = common.points(m, seed=401)
    P, _, _, _ = common.points(n * m, seed=402)
    Q[128 * 2:128 * 3] = 0   # an infinity G2 point in the key
    P[64 * 7:64 * 8] = 0     # and infinity G1 operands (odd and even numbers of live lines per chunk)
    P[64 * 12:64 * 13] = 0
    P[64 * 31:64 * 32] = 0
    out = np.zeros(384 * n, np.uint8)
    emu.emu_multi_pair_lines(vp(P), vp(Q), sz(n), sz(m), vp(out))
    ref = port.multi_pair_batch(P, np.tile(Q, n), n, m)
    assert (out == ref).all()


def test_xi_multiplication_fast_path(emu):
    """(9+u) * a through the shift / small-quotient reduction (tower.cuh fp_mul9_add) against the oracle: random values
    and the boundary cases of the quotient estimate (components 0, 1, p-1, p-2, values around k*p/9)."""
    rng = o.SplitMix64(404)
    vals = [0, 1, 2, o.P - 1, o.P - 2, o.P // 2, o.P // 9, o.P // 9 + 1, 2 * o.P // 9, (o.P - 1) // 3, 8 * o.P // 9 + 1]
    pairs = [(x, y) for x in vals for y in vals] + [(rng.fp(), rng.fp()) for _ in range(4000)]
    n = len(pairs)
    a = np.frombuffer(b"".join(o.fp_to_mont_bytes(x) + o.fp_to_mont_bytes(y) for x, y in pairs), dtype=np.uint8).copy()
    z = np.zeros(64 * n, dtype=np.uint8)
    emu.emu_fp2_mul_xi(vp(a), sz(n), vp(z))
    for i, (x, y) in enumerate(pairs):
        e = o.fp2_mul_xi((x, y))
        assert z[64 * i:64 * i + 64].tobytes() == o.fp_to_mont_bytes(e[0]) + o.fp_to_mont_bytes(e[1]), (x, y)


def test_warp_vm_programs_on_host(emu):
    """The warp-VM (one warp per pairing, Fp-level ops: wvmgen.py programs run by wvm.cuh's exec_op) with lock-step
    round semantics: Miller-only + final-exp-only == pairing == oracle, final exponentiation of an arbitrary Fp12."""
    n = 3
    P, Q, _, _ = common.points(n, seed=321)
    ref = port.pair_batch(P, Q, n)
    out = np.zeros(384 * n, np.uint8)
    emu.emu_wvm(vp(P), vp(Q), sz(n), 1, vp(out))
    assert (out == ref).all()
    ml = np.zeros(384 * n, np.uint8)
    emu.emu_wvm(vp(P), vp(Q), sz(n), 0, vp(ml))
    fe = np.zeros(384 * n, np.uint8)
    emu.emu_wvm(vp(ml), None, sz(n), 2, vp(fe))
    assert (fe == ref).all()
    assert (port.final_exp_batch(ml, n) == ref).all()  # the Miller value itself is a valid input of the oracle's FE
    rng = o.SplitMix64(78)
    x = np.frombuffer(b"".join(o.fp_to_mont_bytes(rng.fp()) for _ in range(12)), dtype=np.uint8).copy()
    emu.emu_wvm(vp(x), None, sz(1), 2, vp(fe))
    assert (fe[:384] == port.final_exp_batch(x, 1)).all()


def test_warp_vm_two_pair_miller_program_on_host(emu):
    """wvmgen's "miller2" program (two Miller loops with shared squarings, one warp: the BLS verification shape) run by
    wvm.cuh's exec_op on the host == the product of the two single-pair Miller values bit for bit, and its final
    exponentiation == the oracle's 2-pair product."""
    n = 2
    P, Q, _, _ = common.points(2 * n, seed=322)
    out = np.zeros(384 * n, np.uint8)
    emu.emu_wvm_miller2(vp(P), vp(Q), sz(n), vp(out))
    single = np.zeros(384 * 2 * n, np.uint8)
    emu.emu_wvm(vp(P), vp(Q), sz(2 * n), 0, vp(single))
    want = port.gt_mul_batch(single.reshape(n, 2, 384)[:, 0].copy().reshape(-1), single.reshape(n, 2, 384)[:, 1].copy().reshape(-1), n)
    assert (out == want).all()
    assert (port.final_exp_batch(out, n) == port.multi_pair_batch(P, Q, n, 2)).all()


def test_warp_vm_lin_reduction_bounds(emu):
    """lin_reduce accepts any 9-limb v < 256 p: multiples of p and their neighbours up to the bound, and random values."""
    rng = o.SplitMix64(79)
    vals = []
    for q in list(range(0, 256, 5)) + [1, 2, 199, 200, 254, 255]:
        for d in (-1, 0, 1, o.P - 1, rng.fp()):
            v = q * o.P + d
            if 0 <= v < 256 * o.P:
                vals.append(v)
    vals += [rng.u256() * 256 % (256 * o.P) for _ in range(5000)]
    buf = np.frombuffer(b"".join(v.to_bytes(36, "little") for v in vals), dtype=np.uint8).copy()
    out = np.zeros(32 * len(vals), np.uint8)
    emu.emu_wvm_lin_reduce(vp(buf), sz(len(vals)), vp(out))
    got = [int.from_bytes(out[32 * i:32 * i + 32].tobytes(), "little") for i in range(len(vals))]
    assert got == [v % o.P for v in vals]


def test_warp_vm_python_evaluator_matches_oracle():
    """wvmgen's own integer evaluator on the scheduled programs (independent of the C++ interpreter)."""
    import sys

    sys.path.insert(0, CSRC)
    import wvmgen as w

    rng = o.SplitMix64(6)
    Pt, Qt = o.g1_mul(o.G1_GEN, rng.scalar()), o.g2_mul(o.G2_GEN, rng.scalar())
    words, meta = w.build("miller")
    init = {s: v for s, v in meta["consts"]}
    ins = meta["in_slots"]
    init.update({ins[0]: Pt[0], ins[1]: Pt[1], ins[2]: Qt[0][0], ins[3]: Qt[0][1], ins[4]: Qt[1][0], ins[5]: Qt[1][1]})
    slots = w.evaluate(words, meta["nslots"], init)
    ml = [slots[s] for s in meta["out_slots"]]
    words, meta = w.build("finalexp")
    init = {s: v for s, v in meta["consts"]}
    init.update({s: v for s, v in zip(meta["in_slots"], ml)})
    slots = w.evaluate(words, meta["nslots"], init)
    got = [slots[s] for s in meta["out_slots"]]
    e = o.pair([Pt], [Qt])
    want = [e[c][b][a] for c in (0, 1) for b in (0, 1, 2) for a in (0, 1)]
    assert got == want


def test_warp_vm_lin_op_and_inversion(emu):
    """The carry-free LIN accumulator (negative terms as complements, sum |c| <= 120) and the binary-Euclid inversion
    against Python integers, including the extreme coefficient patterns and inv(0) = 0."""
    rng = o.SplitMix64(80)
    R = 1 << 256
    cases = [([o.P - 1] * 8, [15] * 8), ([o.P - 1] * 8, [-15] * 8), ([0] * 4, [-30] * 4), ([1, o.P - 1], [31, -31]), ([5], [1]), ([7], [-1])]
    for _ in range(300):
        k = 1 + rng.next() % 8
        cs = []
        budget = 120
        for _ in range(k):
            c = int(rng.next() % 63) - 31
            c = max(-budget, min(budget, c)) if budget else 0
            budget -= abs(c)
            cs.append(c)
        cases.append(([rng.fp() for _ in range(k)], cs))
    for vals, cs in cases:
        buf = np.frombuffer(b"".join(v.to_bytes(32, "little") for v in vals), dtype=np.uint8).copy()
        carr = (ctypes.c_int * len(cs))(*cs)
        out = np.zeros(32, np.uint8)
        emu.emu_wvm_lin(vp(buf), carr, sz(len(cs)), vp(out))
        assert int.from_bytes(out.tobytes(), "little") == sum(c * v for c, v in zip(cs, vals)) % o.P, (vals, cs)
    xs = [0, 1, 2, o.P - 1, (1 << 255) % o.P] + [rng.fp() for _ in range(200)]
    buf = np.frombuffer(b"".join(x.to_bytes(32, "little") for x in xs), dtype=np.uint8).copy()
    out = np.zeros(32 * len(xs), np.uint8)
    emu.emu_wvm_inv(vp(buf), sz(len(xs)), vp(out))
    for i, x in enumerate(xs):  # Montgomery in, Montgomery out: (a R) -> a^-1 R
        want = 0 if x == 0 else pow(x * pow(R, -1, o.P) % o.P, -1, o.P) * R % o.P
        assert int.from_bytes(out[32 * i:32 * i + 32].tobytes(), "little") == want
