"""Small end-to-end run for compute-sanitizer / first-contact debugging on the GPU box."""
import sys
import numpy as np
sys.path.insert(0, ".")
from hysortk_b200 import capi, synth
from oracle import pyoracle as po

k, m, ext = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (31, 17, 0)
rs = synth.sample_mixed(20000, 400, [k - 1, k, 97, 150, 263, 2000], 0.01, seed=7)
exp = po.kmer_count(rs.packed, rs.readlens, k, m, 2, 50, ext, via_supermers=False)
bpr = int(sys.argv[4]) if len(sys.argv) > 4 else 32   # 0: bins sized from the input (small bins, de-duplication)
with capi.Context(k, m, 2, 50, ext, buckets_per_rank=bpr) as ctx:
    r = ctx.count(rs.packed, rs.readlens)
    print("stats", r["stats"])
    got = po.canonicalize(k, r["words"], r["cnt"], r.get("occ_off"), r.get("pos"), r.get("rid"))
    print("kept", got.n, "expected", exp.n)
    po.assert_equal(got, exp, "sanity")
print("OK")
