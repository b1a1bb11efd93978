// Tower-VM interpreter: K lanes of a warp cooperate on one pairing; every tower value lives in a 64-byte
// slot of shared memory (hot) or of an L2-resident global scratch (cold, slot id >= COLD_BASE).  The
// program (vmgen.py) is straight-line: ROUNDS of K same-opcode Fp2 micro-ops, one per lane, separated by
// __syncwarp().  Why: ncu on the one-thread-per-pairing kernel (profiles/r1) showed the path bound by its
// 7 KB/thread local stack (3 TB/s DRAM traffic) and, before that, by instruction-cache misses; here the
// whole state is on-chip (2.3 KB per pairing) and the interpreter + leaves are ~25 KB of code.
//
// Replaces (reference side): the same gnark calls as pairing.cuh (bn254.Pair / MillerLoop /
// FinalExponentiation; access/tree/access_tree_node.go:106,110, cpabe/bsw07/bsw07_cpabe.go:184).
#pragma once
#include "pairing.cuh"
#include "vm_prog_meta.cuh"

namespace bn254 { namespace vm {

enum : int { OP_NOP, OP_MUL, OP_SQR, OP_ADD, OP_SUB, OP_SUB2, OP_DBL, OP_NEG, OP_CONJ, OP_MULXI, OP_HALF, OP_MULFP,
             OP_MULC, OP_MULCFP, OP_MOV, OP_ADDXI, OP_TRIPLE, OP_INV, OP_LDC, OP_SUBXI };
constexpr int COLD_BASE = 160;
constexpr int SLOT_NONE = 0xFF;

// Per-lane view of one pairing's slot file.
//   hot : uint4 index (c * nslots + s) * hot_stride + pid          (shared memory; hot_stride odd)
//   cold: uint4 index ((s - COLD_BASE) * 4 + c) * cold_stride + gpid (global scratch, coalesced over pairings)
struct SlotFile {
  uint4* hot;
  uint4* cold;
  int nslots, hot_stride, pid;
  int cold_stride, gpid;
};

#if defined(__CUDACC__)
BN_D Fp2 ld_slot(const SlotFile& f, int s) {
  Fp2 v;
  uint4* d = reinterpret_cast<uint4*>(&v);
  if (s >= COLD_BASE) {
    const uint4* p = f.cold + (size_t)(s - COLD_BASE) * 4 * f.cold_stride + f.gpid;
#pragma unroll
    for (int c = 0; c < 4; c++) d[c] = p[(size_t)c * f.cold_stride];
  } else {
    const uint4* p = f.hot + s * f.hot_stride + f.pid;
    int cs = f.nslots * f.hot_stride;
#pragma unroll
    for (int c = 0; c < 4; c++) d[c] = p[c * cs];
  }
  return v;
}
BN_D void st_slot(const SlotFile& f, int s, const Fp2& v) {
  const uint4* d = reinterpret_cast<const uint4*>(&v);
  if (s >= COLD_BASE) {
    uint4* p = f.cold + (size_t)(s - COLD_BASE) * 4 * f.cold_stride + f.gpid;
#pragma unroll
    for (int c = 0; c < 4; c++) p[(size_t)c * f.cold_stride] = d[c];
  } else {
    uint4* p = f.hot + s * f.hot_stride + f.pid;
    int cs = f.nslots * f.hot_stride;
#pragma unroll
    for (int c = 0; c < 4; c++) p[c * cs] = d[c];
  }
}
BN_D void vm_sync() { __syncwarp(); }
#else
// host emulation: a slot file is a plain array of Fp2 (hot and cold in one vector of 256)
BN_D Fp2 ld_slot(const SlotFile& f, int s) { return reinterpret_cast<const Fp2*>(f.hot)[s]; }
BN_D void st_slot(const SlotFile& f, int s, const Fp2& v) { reinterpret_cast<Fp2*>(f.hot)[s] = v; }
BN_D void vm_sync() {}
#endif

BN_HD Fp2 fp2_triple_i(const Fp2& a) { return fp2_add_i(fp2_dbl_i(a), a); }

// One micro-op: returns true and the value to store in `out` when the op writes a slot.
BN_HD bool exec_op(const SlotFile& f, uint64_t w, int& dst, Fp2& out) {
  int op = (int)(w & 0xFF);
  if (op == OP_NOP) return false;
  dst = (int)((w >> 8) & 0xFF);
  int a = (int)((w >> 16) & 0xFF), a2 = (int)((w >> 24) & 0xFF), b = (int)((w >> 32) & 0xFF), b2 = (int)((w >> 40) & 0xFF);
  int imm = (int)((w >> 48) & 0xFF);
  Fp2 A, B;
  if (a != SLOT_NONE) A = ld_slot(f, a);
  bool do_mul = false;  // MUL and MULC share ONE copy of the product code after the switch
  switch (op) {
    case OP_MUL:
      if (a2 != SLOT_NONE) A = fp2_add_i(A, ld_slot(f, a2));
      B = ld_slot(f, b);
      if (b2 != SLOT_NONE) B = fp2_add_i(B, ld_slot(f, b2));
      do_mul = true;
      break;
    case OP_SQR:
      if (a2 != SLOT_NONE) A = fp2_add_i(A, ld_slot(f, a2));
      out = fp2_sqr_inl(A);
      break;
    case OP_ADD: out = fp2_add_i(A, ld_slot(f, b)); break;
    case OP_SUB: out = fp2_sub_i(A, ld_slot(f, b)); break;
    case OP_SUB2: out = fp2_sub_i(fp2_sub_i(A, ld_slot(f, b)), ld_slot(f, b2)); break;
    case OP_DBL: out = fp2_dbl_i(A); break;
    case OP_TRIPLE: out = fp2_triple_i(A); break;
    case OP_NEG: out = fp2_neg_i(A); break;
    case OP_CONJ: out = fp2_conj_i(A); break;
    case OP_MULXI: out = fp2_mul_xi_i(A); break;
    case OP_ADDXI: out = fp2_add_i(A, fp2_mul_xi_i(ld_slot(f, b))); break;
    case OP_SUBXI: out = fp2_sub_i(A, fp2_mul_xi_i(ld_slot(f, b))); break;
    case OP_HALF: out = fp2_half_i(A); break;
    case OP_MULFP: {
      B = ld_slot(f, b);
      out = fp2_mul_fp_i(A, (imm & 1) ? B.a1 : B.a0);
      break;
    }
    case OP_MULC: B = VM_CONST2[imm]; do_mul = true; break;
    case OP_MULCFP: { Fp c = VM_CONSTFP[imm]; out = fp2_mul_fp_i(A, c); break; }
    case OP_MOV: out = A; break;
    case OP_LDC: out = VM_CONST2[imm]; break;
    case OP_INV: fp2_inv(out, A); break;
    default: return false;
  }
  if (do_mul) {
    out = fp2_mul_best(A, B);
    if (op == OP_MUL) {  // fused Karatsuba recombination: (a+a2)(b+b2) - c - e
      int c = imm, e = (int)((w >> 56) & 0xFF);
      if (c != SLOT_NONE) out = fp2_sub_i(out, ld_slot(f, c));
      if (e != SLOT_NONE) out = fp2_sub_i(out, ld_slot(f, e));
    }
  }
  return true;
}

// Run `rounds` rounds of a K-wide program for the pairing this lane group owns.  `j` = lane index inside
// the group (j >= K or an out-of-range pairing => the lane only takes part in the warp syncs).
template <int K>
BN_HD void run(const SlotFile& f, const uint64_t* __restrict__ prog, int rounds, int j, bool active) {
  uint64_t w = active ? prog[j] : 0;
  for (int r = 0; r < rounds; r++) {
    uint64_t wn = (active && r + 1 < rounds) ? prog[(size_t)(r + 1) * K + j] : 0;  // prefetch next round's op
    int dst = 0;
    Fp2 out;
    bool st = exec_op(f, w, dst, out);
    if (st) st_slot(f, dst, out);
    vm_sync();
    w = wn;
  }
}

} }  // namespace bn254::vm
