#!/usr/bin/env python3
"""Per-source-line hot spots of an .ncu-rep captured with --import-source on (-lineinfo build):
   python tools/ncu_lines.py file.ncu-rep [top]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
fname = "?"; hdr = None; data = []
for r in csv.reader(out.splitlines()):
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; ci = r.index("Instructions Executed"); si = r.index("# Samples"); continue
    if hdr and r[0] != "" and r[0].isdigit():
        try: data.append((fname, int(r[0]), r[1], int(r[ci]), int(r[si])))
        except ValueError: pass
tot = sum(d[3] for d in data) or 1; ts = sum(d[4] for d in data) or 1
print(f"total warp instructions {tot}, samples {ts}")
for d in sorted(data, key=lambda d: -d[4])[:top]:
    print(f"{d[0]:>12}:{d[1]:<5d} {100*d[3]/tot:5.1f}% inst {100*d[4]/ts:5.1f}% samp  {d[2].strip()[:100]}")
