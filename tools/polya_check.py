"""Heavy-hitter input (the reference pre-counts tasks far above the average, kmerops.cpp:1157-1199):
10^9 bases of poly-A (100 000 reads of 10 kbp) among 200 Mbp of ordinary reads.  One minimizer bin holds ~10^9 k-mers and
~3.3 * 10^7 supermer slots.  Checks: the call completes, the k-mer total is exact, the result equals the result of the
ordinary reads alone (AAA...A is counted ~10^9 times and dropped by UPPER).   Usage: python tools/polya_check.py [Gbases]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hysortk_b200 import capi, synth  # noqa: E402

K, M = 31, 17


def main():
    gb = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    L = 10_000
    rs = synth.sample_fixed(10_000_000, 20.0, L, 0.01, seed=5)
    npoly = int(gb * 1e9) // L
    nb = L // 4
    packed = np.concatenate([rs.packed, np.zeros(npoly * nb, dtype=np.uint8)])
    lens = np.concatenate([rs.readlens, np.full(npoly, L, dtype=np.uint64)])
    N = rs.num_kmers(K) + npoly * (L - K + 1)
    with capi.Context(K, M, 2, 50) as ctx:
        base = ctx.count(rs.packed, rs.readlens)
        t0 = time.time()
        r = ctx.count(packed, lens)
        dt = time.time() - t0
        st = r["stats"]
        assert st["n_kmers_local"] == N, (st["n_kmers_local"], N)
        a = np.sort(base["words"][:, 0] * np.uint64(64) + base["cnt"].astype(np.uint64))
        b = np.sort(r["words"][:, 0] * np.uint64(64) + r["cnt"].astype(np.uint64))
        assert np.array_equal(a, b), "result differs from the result without the poly-A reads"
        print(f"poly-A {gb} Gbases + {rs.nbases / 1e6:.0f} Mbp reads: N = {N} k-mers, kept {r['n_kept']}, overflow bins {st['n_overflow_bins']}, "
              f"{dt * 1e3:.0f} ms (extract {st['ms_extract']:.1f} ms, bins {st['ms_bins']:.1f} ms)", flush=True)
    print("POLYA_CHECK_OK", flush=True)


if __name__ == "__main__":
    main()
