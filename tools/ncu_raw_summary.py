#!/usr/bin/env python
"""Per-kernel summary of an .ncu-rep (ncu --set full): duration, DRAM bytes, throughput %, occupancy, top stalls.
usage: python tools/ncu_raw_summary.py file.ncu-rep"""
import csv
import io
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "rdMB"), ("dram__bytes_write.sum", "wrMB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("launch__registers_per_thread", "regs"),
        ("launch__occupancy_limit_registers", "limR"), ("launch__occupancy_limit_shared_mem", "limS"),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64%"),
        ("smsp__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64%"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64c%"),
        ("lts__t_sector_hit_rate.pct", "l2hit%"), ("l1tex__t_sector_hit_rate.pct", "l1hit%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%")]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, units = rows[0], rows[1]
    ki = h.index("Kernel Name")
    stall = [(i, n) for i, n in enumerate(h) if n.startswith("smsp__average_warps_issue_stalled") and n.endswith("_per_issue_active.ratio")]
    if not stall:
        stall = [(i, n) for i, n in enumerate(h) if n.startswith("smsp__average_warp_latency_issue_stalled") or n.startswith("smsp__average_warps_issue_stalled")]
    for r in rows[2:]:
        name = r[ki].split("(")[0][:60]
        parts = []
        for m, lab in WANT:
            if m in h:
                v = r[h.index(m)]
                u = units[h.index(m)]
                try:
                    fv = float(v.replace(",", ""))
                    if lab in ("rdMB", "wrMB"):
                        scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}.get(u, 1)
                        fv *= scale
                    if lab == "us":
                        scale = {"ns": 1e-3, "us": 1, "ms": 1e3, "usecond": 1, "nsecond": 1e-3, "msecond": 1e3}.get(u, 1)
                        fv *= scale
                    parts.append(f"{lab}={fv:.1f}")
                except ValueError:
                    parts.append(f"{lab}={v}")
        st = []
        for i, n in stall:
            try:
                st.append((float(r[i]), n.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
            except ValueError:
                pass
        st.sort(reverse=True)
        print(name)
        print("   " + " ".join(parts))
        print("   stalls: " + ", ".join(f"{n}={v:.2f}" for v, n in st[:5]))


if __name__ == "__main__":
    main(sys.argv[1])
