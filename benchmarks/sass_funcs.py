#!/usr/bin/env python
"""Static SASS footprint of one kernel of the built library, split at CALL targets:
   python benchmarks/sass_funcs.py <lib.so> <kernel-name-substring>
Used to keep the hot loop inside the SM instruction caches (profiles/r1: no_instruction stalls)."""
import re
import subprocess
import sys
from collections import Counter

lib, want = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(txt) if "Function :" in l and want in l)
end = next((i for i in range(start + 1, len(txt)) if "Function :" in txt[i]), len(txt))
ins = []
for l in txt[start:end]:
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
print("kernel code bytes: %d" % (ins[-1][0] + 16))
tg = sorted({int(m.group(1), 16) for a, t in ins for m in [re.search(r"CALL\.REL\.NOINC (0x[0-9a-f]+)", t)] if m})
ncalls = Counter(re.search(r"CALL\.REL\.NOINC (0x[0-9a-f]+)", t).group(1) for a, t in ins if "CALL.REL" in t)
tg.append(ins[-1][0] + 16)
for i, t in enumerate(tg[:-1]):
    body = [x for a, x in ins if t <= a < tg[i + 1]]
    c = Counter((x.split()[1] if x.startswith("@") else x.split()[0]) for x in body)
    calls = Counter(re.search(r"CALL\.REL\.NOINC (0x[0-9a-f]+)", x).group(1) for x in body if "CALL.REL" in x)
    print("%7s %5d instrs  call-sites %3d  %s  -> %s" % (hex(t), len(body), ncalls[hex(t)], dict(c.most_common(6)), dict(calls.most_common(5))))
