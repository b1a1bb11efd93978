#!/usr/bin/env python
"""Attribute the executed instructions and stall samples of one kernel in an ncu report to CUDA source lines:
   python benchmarks/ncu_sass_by_line.py <report.ncu-rep> <object.o> <mangled-kernel-substring> [top]
Joins `ncu --page source --print-source sass --csv` (per-instruction counters, needs --import-source on / -lineinfo) with
`nvdisasm --print-line-info` of the same object by instruction offset."""
import csv
import glob
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import Counter

rep, obj, want = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=td, check=True, capture_output=True)
    cubin = glob.glob(os.path.join(td, "*.cubin"))[0]
    dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(dis) if l.startswith("\t.section\t.text.") and want in l)
end = next(i for i in range(start + 1, len(dis)) if dis[i].startswith("\t.section"))
cur, by_off = None, {}
for l in dis[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
    if m:
        by_off[int(m.group(1), 16)] = (cur, m.group(2).strip())
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
h = next(i for i, r in enumerate(rows) if "# Samples" in r)
hdr, data = rows[h], [r for r in rows[h + 1:] if len(r) == len(rows[h])]
ix = {n: i for i, n in enumerate(hdr)}
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
base = int(data[0][0], 16)
ex, sm, agg, ops_e, ops_s = Counter(), Counter(), Counter(), Counter(), Counter()
for r in data:
    e, s = int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]])
    loc, ins = by_off.get(int(r[0], 16) - base, (("?", 0), "?"))
    ex[loc] += e
    sm[loc] += s
    for n in stalls:
        agg[n] += int(r[ix[n]])
    w = ins.split()
    op = (w[1] if w[0].startswith("@") else w[0]) if w else "?"
    op = op.split(".")[0] + (".WIDE" if "WIDE" in op else "")
    ops_e[op] += e
    ops_s[op] += s
te, ts = sum(ex.values()), sum(sm.values())
print("kernel %s: %d warp instructions executed, %d stall samples" % (rows[0][1] if rows and len(rows[0]) > 1 else want, te, ts))
print("stall reasons (share of samples): " + ", ".join("%s %.1f%%" % (n[6:], 100.0 * v / ts) for n, v in agg.most_common(8)))
print("by opcode (executed %, samples %): " + ", ".join("%s %.1f/%.1f" % (o, 100.0 * e / te, 100.0 * ops_s[o] / ts) for o, e in ops_e.most_common(10)))
cache = {}
for loc, e in ex.most_common(top):
    f, n = loc
    if f not in cache:
        p = glob.glob(os.path.join(ROOT, "gopairingbasedcryptography_b200", "csrc", f))
        cache[f] = open(p[0]).read().split("\n") if p else []
    text = cache[f][n - 1].strip()[:100] if 0 < n <= len(cache[f]) else ""
    print("%5.1f%% executed %5.1f%% samples  %s:%d  %s" % (100.0 * e / te, 100.0 * sm[loc] / ts, f, n, text))
