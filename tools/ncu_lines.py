#!/usr/bin/env python3
"""Per-source-line instruction / stall-sample shares from an ncu report (needs -lineinfo and --import-source on).
usage: ncu_lines.py report.ncu-rep [top]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
data = []; fname = ""
for r in rows:
    if len(r) == 2 and r[0] == "File Name": fname = r[1].split("/")[-1]; continue
    if len(r) < 10 or r[0] in ("Line No", ""): continue
    try: data.append((int(r[7]), int(r[6]), fname, r[0], r[1].strip()[:100]))
    except ValueError: pass
ti = sum(d[0] for d in data) or 1; ts = sum(d[1] for d in data) or 1
print(f"total warp-instructions {ti}, samples {ts}")
for d in sorted(data, key=lambda x: -x[1])[:top]:
    print(f"{d[0]/ti*100:5.1f}% inst {d[1]/ts*100:5.1f}% samples  {d[2]}:{d[3]}  {d[4]}")
