#!/usr/bin/env python3
"""Summarise an ncu report per CUDA source line: ncu -i rep --page source --csv --print-source cuda,sass
usage: tools/ncu_lines.py <rep> <kernel-regex> [top]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
fname = None; hdr = None; agg = {}
tot_inst = 0; tot_samp = 0
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < 8: continue
    if r[2] != "-": continue       # only source-line summary rows (Address == '-')
    try:
        line = int(r[0]); samp = int(r[6] or 0); inst = int(r[7] or 0)
    except ValueError: continue
    key = (fname, line, r[1].strip()[:90])
    a = agg.setdefault(key, [0, 0]); a[0] += inst; a[1] += samp
    tot_inst += inst; tot_samp += samp
print(f"total warp-inst {tot_inst}  samples {tot_samp}")
for (f, l, s), (i, sm) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{100*sm/max(tot_samp,1):5.1f}% samp {100*i/max(tot_inst,1):5.1f}% inst  {f}:{l}  {s}")
