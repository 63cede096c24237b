#!/usr/bin/env python
"""Hot SASS of one kernel from a .ncu-rep: total instructions executed, and the instructions with the most stall samples.
usage: ncu_sass_hot.py report.ncu-rep kernel-regex [top]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", f"regex:{kern}"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ia, isrc, isamp, iexec = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
body = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break          # first matching launch only
    if len(r) == len(hdr) and r[0] != "Address":
        body.append(r)
tot_exec = sum(int(r[iexec]) for r in body); tot_samp = sum(int(r[isamp]) for r in body)
print(f"{rows[0][1]}: {len(body)} SASS instructions, {tot_exec} warp-instructions executed, {tot_samp} samples")
cls = {}
for r in body:
    op = r[isrc].split()[0] if not r[isrc].strip().startswith("@") else r[isrc].split()[1]
    op = op.split(".")[0]
    cls[op] = cls.get(op, 0) + int(r[iexec])
print("executed by opcode:", ", ".join(f"{k}={v / tot_exec:.1%}" for k, v in sorted(cls.items(), key=lambda kv: -kv[1])[:14]))
for i, r in sorted(enumerate(body), key=lambda ir: -int(ir[1][isamp]))[:top]:
    print(f"{i:5d} {int(r[isamp]) / max(tot_samp, 1):6.1%} exec={int(r[iexec]):>10d}  {r[isrc].strip()[:110]}")
