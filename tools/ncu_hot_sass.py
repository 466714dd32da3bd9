#!/usr/bin/env python
"""Top stall-sample SASS instructions of one kernel in an .ncu-rep (needs --import-source on / -lineinfo).
usage: python tools/ncu_hot_sass.py file.ncu-rep kernel_regex [launch_skip] [top]"""
import csv, io, subprocess, sys
path, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--kernel-name", "regex:" + rx, "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
print(rows[0][1][:150])
h = rows[1]
iS, iN, iE = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
data = []
for n, r in enumerate(rows[2:]):
    try:
        data.append((int(r[iN]), n, r[iS].strip(), int(r[iE])))
    except (ValueError, IndexError):
        pass
tot = sum(d[0] for d in data)
tinst = sum(d[3] for d in data)
print(f"instructions in kernel: {len(data)}, total samples {tot}, executed warp-instr {tinst}")
for s, n, src, e in sorted(data, reverse=True)[:top]:
    print(f"{100*s/tot:5.1f}%  #{n:5d}  exec={e:9d}  {src[:110]}")
# histogram by opcode
ops = {}
for s, n, src, e in data:
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    op = op.split(".")[0]
    ops.setdefault(op, [0, 0]); ops[op][0] += e; ops[op][1] += s
print("opcode mix (executed warp-instr, % of total | stall samples %):")
for op, (e, s) in sorted(ops.items(), key=lambda kv: -kv[1][0])[:18]:
    print(f"   {op:10s} {100*e/tinst:5.1f}%   {100*s/tot:5.1f}%")
