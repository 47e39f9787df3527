#!/usr/bin/env python3
"""Per-source-line instruction / stall-sample shares from an ncu report (needs -lineinfo and --import-source on).
usage: ncu_lines.py report.ncu-rep [top] [kernel-substring]   (a report with several kernels: pick one by substring)"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
want = sys.argv[3] if len(sys.argv) > 3 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; fname = ""; func = ""; data = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if len(r) == 2 and r[0] == "Function Name": func = r[1]; continue
    if len(r) > 10 and r[0] == "Line No": hdr = r; i_s = hdr.index("# Samples"); i_e = hdr.index("Instructions Executed"); continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit() or want not in func: continue
    try: data.append((int(r[i_e]), int(r[i_s]), fname, r[0], r[1].strip()[:110]))
    except ValueError: pass
ti = sum(d[0] for d in data) or 1; ts = sum(d[1] for d in data) or 1
print(f"total warp-instructions {ti}, samples {ts}")
for d in sorted(data, key=lambda d: -d[1])[:top]:
    print(f"{d[0] / ti * 100:5.1f}% inst {d[1] / ts * 100:5.1f}% samples  {d[2]}:{d[3]}  {d[4]}")
