#!/usr/bin/env python
"""ncu --metrics CSV of benchmarks/profile_driver_r2.py -> the per-family table and executed_imad_per_unit.json:
   python benchmarks/ncu_families_summary.py <metrics.csv> <units.json> <out_dir>"""
import csv
import json
import sys
from collections import OrderedDict

M = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
     "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
     "sm__inst_executed_pipe_fmaheavy.sum", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum"]
rows = OrderedDict()
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    k = (int(r["ID"]), r["Kernel Name"])
    v = r["Metric Value"].replace(",", "")
    rows.setdefault(k, {})[r["Metric Name"]] = float(v) if v not in ("", "n/a") else float("nan")
units = json.load(open(sys.argv[2]))
out = ["kernel%sgrid   time ms  regs  warps%%  issue%%  fmaheavy%%   DRAM MB   fmaheavy thread-instr   units   IMAD-class/unit  DRAM B/unit" % (" " * 40)]
imad = {}
NAMES = {"k_subset_sum_tab": "_subset_sum_tab", "k_gt_mul<0>": "_gt_mul", "k_pair": "pair", "k_check2_fixed_g1": "bls_verify", "k_scalar_mul<bn254::G1Jac": "g1_var", "k_scalar_mul_g2_gls": "g2_var",
         "k_fixed_mul<G1": "g1_fixed", "k_fixed_mul<G2": "g2_fixed", "k_gt_exp<0>": "gt_exp", "k_gt_exp<1>": "gt_cyclo_exp", "k_gt_fixed_exp": "gt_fixed_exp",
         "k_miller_lines": "bsw07_decrypt_policy_lines"}
def short_name(name):
    short = name.replace("void <unnamed>::", "").replace("void bn254::<unnamed>::", "").split("(")[0]
    return short.replace("bn254::", "")


# a unit count belongs to the LARGEST launch of its kernel (table builds launch the same kernels on small grids)
owner = {}
for u in units:
    key = u["kernel"].replace("bn254::", "")
    cands = [(m.get("launch__grid_size", 0), k) for k, m in rows.items() if key in short_name(k[1]) and k not in owner]
    if cands:
        owner[max(cands)[1]] = u
for (i, name), m in rows.items():
    short = short_name(name)
    u = owner.get((i, name))
    t = m.get("gpu__time_duration.sum", 0) / 1e6
    fm = m.get("sm__inst_executed_pipe_fmaheavy.sum", float("nan")) * 32
    dram = m.get("dram__bytes_read.sum", 0) + m.get("dram__bytes_write.sum", 0)
    extra = ""
    if u is not None:
        extra = "%9d %14.4g %12.4g" % (u["units"], fm / u["units"], dram / u["units"])
        for k_, row in NAMES.items():
            if k_.replace("bn254::", "") in short:
                imad[row] = fm / u["units"]
    out.append("%-44s %5d %9.3f %5d %7.1f %7.1f %9.1f %10.1f %18.4g %s" % (
        short[:44], m.get("launch__grid_size", 0), t, m.get("launch__registers_per_thread", 0),
        m.get("sm__warps_active.avg.pct_of_peak_sustained_active", 0), m.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0),
        m.get("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", 0), dram / 1e6, fm, extra))
open(sys.argv[3] + "/ncu_kernel_families_summary.txt", "w").write(
    "one launch per kernel family, benchmarks/profile_driver_r2.py; ncu --metrics pass (cold-cache, serialised: the utilisation and per-unit\n"
    "columns are what this file is for).  IMAD-class/unit = sm__inst_executed_pipe_fmaheavy.sum x 32 / units of the launch.\n\n" + "\n".join(out) + "\n")
# composite flows: sums of the measured per-unit counts of the kernels they launch (derived, not captured as one unit)
if all(k in imad for k in ("bsw07_decrypt_policy_lines", "g1_var")):
    imad["bsw07_decrypt_key_lines"] = imad["bsw07_decrypt_policy_lines"] + 200 * imad["g1_var"]  # + [Delta_i]Cy_i, [Delta_i]Cy'_i per ciphertext
if all(k in imad for k in ("g1_fixed", "_subset_sum_tab", "g2_var", "gt_fixed_exp", "_gt_mul")):
    imad["waters05_encrypt"] = imad["g1_fixed"] + imad["_subset_sum_tab"] + imad["g2_var"] + imad["gt_fixed_exp"] + imad["_gt_mul"]
imad = {k: v for k, v in imad.items() if not k.startswith("_")}
json.dump(imad, open(sys.argv[3] + "/executed_imad_per_unit.json", "w"), indent=1)
print("\n".join(out))
