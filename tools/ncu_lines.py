"""Per-CUDA-source-line totals of one kernel of an .ncu-rep (captured with --import-source on, built with
-lineinfo): stall samples and executed warp instructions per (file, line).
Usage: python tools/ncu_lines.py file.ncu-rep kernel_regex [top]"""
import csv
import io
import subprocess
import sys


def main(rep, kernel, top=40):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv",
                          "--kernel-name", "regex:" + kernel], capture_output=True).stdout.decode("utf-8", "replace")
    rows = list(csv.reader(io.StringIO(out)))
    fname, hdr = "?", None
    agg = {}
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdr = r
        elif hdr and len(r) == len(hdr) and r[0].isdigit():
            # a CUDA line is followed by its SASS rows; rows carrying an Address are SASS rows
            try:
                s = int(r[hdr.index("# Samples")] or 0)
                i = int(r[hdr.index("Instructions Executed")] or 0)
            except ValueError:
                continue
            if r[hdr.index("Address")] != "-":
                continue
            key = (fname, int(r[0]))
            a = agg.setdefault(key, [0, 0, r[1]])
            a[0] += s
            a[1] += i
    ts = sum(a[0] for a in agg.values()) or 1
    ti = sum(a[1] for a in agg.values()) or 1
    print(f"kernel {kernel}: total samples {ts}  warp instructions {ti}")
    for (f, ln), (s, i, src) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{f:12s} {ln:4d} samples {100*s/ts:5.1f}% inst {100*i/ti:5.1f}%  {src.strip()[:110]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
