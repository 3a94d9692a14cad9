"""Size-independent properties at the per-GPU share of BASELINE.json configs[2] (30 Gbp over 8 GPUs = 3.75 Gbp per GPU):
reads are generated on the GPU (uniform genome, 10x coverage, 1 % substitutions, both strands), counted through
hsk_count_device, and checked without an oracle:
  * n_kmers_local == sum(max(len - K + 1, 0))
  * L=2,U=50: the result is the same multiset for two different bin counts (order-independent checksums on the GPU),
    entries ascend inside a bin (descents == bins - 1), histogram == bincount(cnt)
Usage: python tools/big_check.py [Gbp, default 3.75]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hysortk_b200 import capi  # noqa: E402

K, M = 31, 17


def make_reads(total_bases: int, coverage: float, read_len: int, err: float, seed: int):
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(seed)
    G = int(total_bases / coverage)
    genome = torch.randint(0, 4, (G,), dtype=torch.uint8, device=dev, generator=g)
    n = total_bases // read_len
    nb = (read_len + 3) // 4
    packed = torch.zeros(n * nb + 64, dtype=torch.uint8, device=dev)
    ar = torch.arange(read_len, device=dev)
    chunk = max(1, (1 << 27) // read_len)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        starts = torch.randint(0, G - read_len + 1, (e - s,), device=dev, generator=g)
        r = genome[starts[:, None] + ar[None, :]]
        mut = torch.rand(r.shape, device=dev, generator=g) < err
        r = torch.where(mut, (r + torch.randint(1, 4, r.shape, dtype=torch.uint8, device=dev, generator=g)) & 3, r)
        flip = torch.rand((e - s,), device=dev, generator=g) < 0.5
        r = torch.where(flip[:, None], (3 - r).flip(1), r)
        pad = nb * 4 - read_len
        if pad:
            r = torch.cat([r, torch.zeros((e - s, pad), dtype=torch.uint8, device=dev)], 1)
        r = r.view(e - s, nb, 4)
        packed[s * nb:e * nb] = ((r[:, :, 0] << 6) | (r[:, :, 1] << 4) | (r[:, :, 2] << 2) | r[:, :, 3]).reshape(-1)
    off = torch.arange(n + 1, dtype=torch.int64, device=dev) * nb
    lens = torch.full((n,), read_len, dtype=torch.int32, device=dev)
    return packed, off, lens, n, n * nb


def main():
    gbp = float(sys.argv[1]) if len(sys.argv) > 1 else 3.75
    t0 = time.time()
    packed, off, lens, n, nbytes = make_reads(int(gbp * 1e9), 10.0, 10_000, 0.01, 7)
    torch.cuda.synchronize()
    N = n * (10_000 - K + 1)
    print(f"generated {n} reads, {nbytes/1e9:.2f} GB packed, N = {N} k-mers in {time.time()-t0:.1f} s", flush=True)
    res = {}
    for tag, env in (("auto bins", None), ("target 6000", "6000")):
        if env:
            os.environ["HSK_TARGET_BIN"] = env
        with capi.Context(K, M, 2, 50) as ctx:
            r = ctx.count_device(packed.data_ptr(), nbytes, off.data_ptr(), lens.data_ptr(), n)
            r = ctx.count_device(packed.data_ptr(), nbytes, off.data_ptr(), lens.data_ptr(), n)
            st = r.stats.as_dict()
            assert st["n_kmers_local"] == N, (st["n_kmers_local"], N)
            host = ctx.fetch()
            w, c = host["words"][:, 0], host["cnt"]
            h = (w * np.uint64(0x9E3779B97F4A7C15)) ^ (w >> np.uint64(29))
            res[tag] = (len(w), int(c.astype(np.uint64).sum()), int((h * c.astype(np.uint64)).sum()))
            assert np.array_equal(host["histogram"], np.bincount(c, minlength=51).astype(np.uint64))
            desc = int((w[1:] < w[:-1]).sum())
            print(f"{tag}: kept {len(w)} sum(cnt) {res[tag][1]} checksum {res[tag][2]:#x} descents {desc} overflow bins {st['n_overflow_bins']} "
                  f"device {st['ms_total']:.1f} ms = {N / st['ms_total'] / 1e6:.1f} G k-mers/s (extract {st['ms_extract']:.1f}, bins {st['ms_bins']:.1f})", flush=True)
        os.environ.pop("HSK_TARGET_BIN", None)
    assert res["auto bins"] == res["target 6000"], res
    print("BIG_CHECK_OK", flush=True)


if __name__ == "__main__":
    main()
