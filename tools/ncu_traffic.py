#!/usr/bin/env python
"""DRAM traffic per kernel family of one RK3 substep from an `ncu --set full` report -> profiles/ncu_traffic_256.json
(read by bench.py as roofline.traffic).  Families are the ones of bench.py; bytes are per launch group of one substep.

usage: python tools/ncu_traffic.py gpurun_out/r2_substep_full.ncu-rep [nsubsteps_in_capture]"""
import csv
import io
import json
import os
import subprocess
import sys

FAMILY = [("k_momtend", "mom_tend"), ("k_closure", "closure"), ("k_fillps", "fillps"), ("k_tderive_integrate", "tderive_integrate"),
          ("k_rfft", "poisson_core"), ("k_zsolve", "poisson_core"), ("k_ztile", "poisson_core"), ("k_solmpj", "poisson_core")]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main(path, nsub):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, units = rows[0], rows[1]
    ki = h.index("Kernel Name")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    fam, per_kernel = {}, {}
    for r in rows[2:]:
        name = r[ki].split("(")[0].split("<")[0].replace("udg::", "").replace("void ", "").strip()
        tot = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = h.index(m)
            tot += float(r[i].replace(",", "")) * scale.get(units[i], 1.0)
        per_kernel.setdefault(name, []).append(tot)
        for pre, f in FAMILY:
            if name.startswith(pre):
                fam[f] = fam.get(f, 0.0) + tot
                break
    fam = {k: v / nsub for k, v in fam.items()}
    head = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    doc = {"source": f"ncu --set full --clock-control none, {os.path.basename(path)}, {nsub} substep(s) of 256^3 on one B200, kernels of commit {head}",
           "unit": "bytes per substep (dram__bytes_read.sum + dram__bytes_write.sum, summed over the family's launches)",
           "families": fam, "kernels": {k: {"launches": len(v), "bytes_per_launch": sum(v) / len(v)} for k, v in per_kernel.items()}}
    json.dump(doc, open(os.path.join(ROOT, "profiles", "ncu_traffic_256.json"), "w"), indent=1)
    print(json.dumps(doc["families"], indent=1))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1)
