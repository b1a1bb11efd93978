#!/usr/bin/env python
"""Summarise an ncu report of one kernel:  python benchmarks/ncu_summary.py report.ncu-rep [kernel-substring]

Prints (1) the raw-page metrics the DESIGN.md tables quote, (2) samples / executed instructions by opcode with the
top stall reasons, (3) the same by device function (split at the CALL targets of the SASS), so a capture brought
back from the GPU box can be read without the GUI.  Output goes under profiles/ when it is worth keeping."""
import csv
import io
import re
import subprocess
import sys
from collections import Counter, defaultdict

RAW = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_elapsed.max",
]
STALLS = ["long_scoreboard", "wait", "math_pipe_throttle", "dispatch_stall", "not_selected", "selected", "branch_resolving",
          "no_instruction", "short_scoreboard", "lg_throttle", "mio_throttle", "barrier", "imc_miss"]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return out


def main():
    rep = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    rows = list(csv.reader(io.StringIO(ncu_csv(rep, "raw"))))
    hdr = rows[0]
    units = rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        if want and want not in name:
            continue
        print("== kernel:", name)
        for m in RAW:
            if m in hdr:
                print("%s [%s] = %s" % (m, units[hdr.index(m)], r[hdr.index(m)]))
        for s in STALLS:
            m = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % s
            if m in hdr:
                print("stall %-20s %s" % (s, r[hdr.index(m)]))
        break
    src = ncu_csv(rep, "source")
    lines = src.split("\n")
    # find the header row of the (first matching) kernel
    start = 0
    for i, l in enumerate(lines):
        if l.startswith('"Kernel Name"') and (not want or want in l):
            start = i + 1
            break
    body = []
    for l in lines[start:]:
        if l.startswith('"Kernel Name"'):
            break
        body.append(l)
    rd = list(csv.reader(io.StringIO("\n".join(body))))
    if not rd:
        return
    h = rd[0]
    ia, isrc, isamp, iex = h.index("Address"), h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    stall_cols = [(c, h.index(c)) for c in h if c.startswith("stall_") and "(Not Issued)" not in c]
    insts = []
    for r in rd[1:]:
        if len(r) <= iex or not r[ia].startswith("0x"):
            continue
        insts.append((int(r[ia], 16), r[isrc].strip(), int(r[isamp] or 0), int(r[iex] or 0), {c: int(r[i] or 0) for c, i in stall_cols}))
    base = insts[0][0]
    tot_s = sum(x[2] for x in insts) or 1
    tot_e = sum(x[3] for x in insts) or 1
    print("\n-- by opcode: total samples %d, warp instructions %d" % (tot_s, tot_e))
    by = defaultdict(lambda: [0, 0, Counter()])
    for a, s, ns, ne, st in insts:
        t = s.split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0] if t else "?"
        by[op][0] += ns
        by[op][1] += ne
        by[op][2].update(st)
    for op, (ns, ne, st) in sorted(by.items(), key=lambda kv: -kv[1][0])[:14]:
        tops = ", ".join("%s %d%%" % (k.replace("stall_", ""), 100 * v // max(1, sum(st.values()))) for k, v in st.most_common(3))
        print("%-8s samples %6.2f%%  executed %6.2f%%  top stalls: %s" % (op, 100.0 * ns / tot_s, 100.0 * ne / tot_e, tops))
    # by function: split at CALL targets
    targets = sorted({int(m.group(1), 16) for a, s, _, _, _ in insts for m in [re.search(r"CALL\.(?:REL|ABS)\.NOINC\s+(0x[0-9a-f]+)", s)] if m})
    targets = [t for t in targets if insts[0][0] <= t <= insts[-1][0]]
    if targets:
        print("\n-- by device function (offset: instrs, samples %, executed %, calls, top stalls)")
        bounds = [base] + targets + [insts[-1][0] + 16]
        fn = []
        for lo, hi in zip(bounds[:-1], bounds[1:]):
            seg = [x for x in insts if lo <= x[0] < hi]
            if not seg:
                continue
            ns = sum(x[2] for x in seg)
            ne = sum(x[3] for x in seg)
            st = Counter()
            for x in seg:
                st.update(x[4])
            calls = seg[0][3]
            fn.append((lo - base, len(seg), ns, ne, calls, st))
        for off, n, ns, ne, calls, st in sorted(fn, key=lambda x: -x[2])[:24]:
            tops = ", ".join("%s %d%%" % (k.replace("stall_", ""), 100 * v // max(1, sum(st.values()))) for k, v in st.most_common(3))
            print("0x%05x: %5d instrs  samples %6.2f%%  executed %6.2f%%  entered %9d  %s" % (off, n, 100.0 * ns / tot_s, 100.0 * ne / tot_e, calls, tops))


if __name__ == "__main__":
    main()
