#!/usr/bin/env python3
"""DRAM traffic per image and stage from an `ncu --set full` capture of tools/profile_run.py, for bench.py's
`roofline.traffic`:  tools/ncu_traffic.py <rep> <images per launch> > profiles/r2_traffic.json
Sums dram__bytes_read.sum + dram__bytes_write.sum over the kernels of each stage (one launch of each) and divides by
the images one launch covered."""
import csv
import json
import subprocess
import sys

STAGE_OF = {"unstuff_count_kernel": "unstuff", "scan_tiles_kernel": "unstuff", "unstuff_scatter_kernel": "unstuff",
            "spec_kernel": "spec", "fix_local_kernel": "fix", "chain_kernel": "fix", "write_kernel": "write",
            "bj_pixels_420_kernel": "pixels"}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    rep, images = sys.argv[1], int(sys.argv[2])
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
    seen, per_stage = set(), {}
    for r in rows[2:]:
        name = r[ik].split("(")[0].replace("<unnamed>::", "").replace("void ", "").strip()
        if name not in STAGE_OF or name in seen:
            continue                      # first captured launch of every kernel
        seen.add(name)
        b = float(r[ir]) * UNIT[units[ir]] + float(r[iw]) * UNIT[units[iw]]
        per_stage[STAGE_OF[name]] = per_stage.get(STAGE_OF[name], 0.0) + b
    out = {"source": f"ncu --set full capture {rep.split('/')[-1]} ({images} images per launch): dram__bytes_read.sum + dram__bytes_write.sum",
           "images_per_launch": images, "kernels": sorted(seen),
           "bytes_per_image": {k: v / images for k, v in per_stage.items()}}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
