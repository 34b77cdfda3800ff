"""Per-source-line share of the thread instructions of a kernel, from an .ncu-rep taken with --import-source on:
    python tools/ncu_lines.py gpurun_out/x.ncu-rep [indices_per_launch] [top]
Uses `ncu --page source --print-source cuda,sass --csv`; rows with a line number carry that line's aggregate."""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
per = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, hdr = None, None
agg = collections.Counter()
src = {}
for r in rows:
    if len(r) >= 2 and r[0] in ("File Path", "File Name"):
        cur = r[1].split("/")[-1]
        continue
    if len(r) > 5 and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[0].isdigit() and cur:
        i = hdr.index("Thread Instructions Executed")
        try:
            v = int(r[i])
        except ValueError:
            v = 0
        if v:
            agg[(cur, int(r[0]))] += v
            src[(cur, int(r[0]))] = r[1].strip()
tot = sum(agg.values())
print(f"# {rep}: {tot} thread instructions, {tot / per:.2f} per unit")
byfile = collections.Counter()
for (f, l), v in agg.items():
    byfile[f] += v
for f, v in byfile.most_common():
    print(f"## {f}: {v / per:.2f} per unit ({100 * v / tot:.1f} %)")
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
    print(f"{f}:{l:<5d} {v / per:8.2f} {100 * v / tot:5.1f}%  {src[(f, l)][:110]}")
