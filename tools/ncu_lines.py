"""Per-CUDA-source-line totals from `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv`:
stall samples and executed warp instructions.  Usage: python tools/ncu_lines.py dump.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = []
for r in rows:
    if len(r) > 8 and r[0].isdigit():
        try:
            out.append((int(r[6] or 0), int(r[7] or 0), int(r[0]), r[1]))
        except ValueError:
            pass
ts = sum(o[0] for o in out) or 1
ti = sum(o[1] for o in out) or 1
print(f"total samples {ts}  warp instructions {ti}")
for s, i, ln, src in sorted(out, reverse=True)[:top]:
    print(f"line {ln:4d} samples {100*s/ts:5.1f}% inst {100*i/ti:5.1f}%  {src.strip()[:120]}")
