"""Per-launch DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) and duration of the kernels of an .ncu-rep,
averaged per kernel name; optionally records one of them in profiles/traffic.json under the key bench.py looks up.
Usage: python tools/ncu_traffic.py file.ncu-rep [kernel_substring key]"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(path, sub=None, key=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ir, iw, it, iname = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum"), hdr.index("Kernel Name")
    agg = collections.OrderedDict()
    for r in rows[2:]:
        b = float(r[ir].replace(",", "")) * UNIT[units[ir]] + float(r[iw].replace(",", "")) * UNIT[units[iw]]
        agg.setdefault(r[iname], []).append((b, float(r[it].replace(",", "")), units[it]))
    for name, v in agg.items():
        b = sum(x[0] for x in v) / len(v)
        print(f"{name[:70]:70s} launches={len(v)} dram_bytes/launch={b/1e6:.1f} MB  time={sum(x[1] for x in v)/len(v):.1f} {v[0][2]}")
        if sub and sub in name and key:
            p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
            d = json.load(open(p)) if os.path.exists(p) else {}
            d[key] = b
            json.dump(d, open(p, "w"), indent=1, sort_keys=True)
            print(f"  -> profiles/traffic.json[{key}] = {b:.0f}")


if __name__ == "__main__":
    main(*sys.argv[1:4])
