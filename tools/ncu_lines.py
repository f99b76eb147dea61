#!/usr/bin/env python
"""Per-source-line view of an .ncu-rep captured with --import-source on: stall samples and executed warp instructions
per CUDA source line (needs -lineinfo).  usage: tools/ncu_lines.py prof.ncu-rep [min_samples]"""
import csv, subprocess, sys
path = sys.argv[1]; thr = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
fname = None; hdr = None; rows = []
for r in csv.reader(out.splitlines()):
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or r[0] in ("", "..."): continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    try:
        rows.append((fname, ln, r[1].strip()[:110], int(r[hdr.index("# Samples")] or 0), int(r[hdr.index("Instructions Executed")] or 0), int(r[hdr.index("Thread Instructions Executed")] or 0)))
    except ValueError:
        pass
ts = sum(x[3] for x in rows) or 1; ti = sum(x[4] for x in rows) or 1
print("total samples %d, warp instructions %d" % (ts, ti))
for f, ln, src, s, ie, te in rows:
    if s >= thr or ie >= ti * 0.01:
        print("%-16s %4d  smp %5d %5.1f%%  inst %9d %5.1f%%  lanes %4.1f | %s" % (f, ln, s, 100.0 * s / ts, ie, 100.0 * ie / ti, te / max(ie, 1), src))
if len(sys.argv) > 3:   # phase table: name=file:lo-hi,...
    print()
    for spec in sys.argv[3].split(","):
        name, rng = spec.split("="); f, lr = rng.split(":"); lo, hi = map(int, lr.split("-"))
        s = sum(x[3] for x in rows if x[0].startswith(f) and lo <= x[1] <= hi); ie = sum(x[4] for x in rows if x[0].startswith(f) and lo <= x[1] <= hi)
        print("%-22s samples %5.1f%%  instructions %5.1f%% (%d)" % (name, 100.0 * s / ts, 100.0 * ie / ti, ie))
