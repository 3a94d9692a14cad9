"""Multi-rank glue above the C ABI (one process per GPU): read partition, NCCL id broadcast, global
ReadId base.  Works with any torch.distributed backend (gloo on CPU for the host logic, nccl on the
GPU box); the data path itself (supermer all-to-all) runs inside libhysortk_b200.so over NCCL.

Mirrors what the reference does with MPI around kmer_count: contiguous read ranges balanced by
bases (reference src/fastaindex.cpp:52-100), MPI_Exscan of the read counts for the ReadId base
(src/kmerops.cpp:65-70), MPI_Allreduce of the histogram (src/hysortk.cpp:104,115).
"""
from __future__ import annotations

import numpy as np

from . import capi


def partition_reads(readlens: np.ndarray, nranks: int) -> np.ndarray:
    """first read of every rank (nranks+1 entries): the reference's greedy contiguous partition —
    keep adding reads until the next one would reach the average number of bases per rank; the last
    rank takes the rest (reference src/fastaindex.cpp:52-100)."""
    lens = np.asarray(readlens, dtype=np.int64)
    n = len(lens)
    avg = float(lens.sum()) / nranks
    first = np.full(nranks + 1, n, dtype=np.int64)
    rid = 0
    for p in range(nranks - 1):
        first[p] = rid
        sofar = 0
        if rid < n:
            while True:
                sofar += int(lens[rid])
                rid += 1
                if not (rid < n and sofar + int(lens[rid]) < avg):
                    break
    first[nranks - 1] = rid
    return first


def shard(packed: np.ndarray, readlens: np.ndarray, first: np.ndarray, rank: int):
    """(packed bytes, read lengths, ReadId base) of one rank's contiguous share."""
    nb = (np.asarray(readlens, dtype=np.uint64) + np.uint64(3)) // np.uint64(4)
    off = np.zeros(len(readlens) + 1, dtype=np.uint64)
    np.cumsum(nb, out=off[1:])
    lo, hi = int(first[rank]), int(first[rank + 1])
    return packed[int(off[lo]):int(off[hi])], readlens[lo:hi], lo


def readid_base(nreads_local: int) -> int:
    """exclusive prefix sum of the per-rank read counts (MPI_Exscan, reference kmerops.cpp:65-70)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return 0
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    mine = torch.tensor([nreads_local], dtype=torch.int64, device=dev)
    allc = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(allc, mine)
    return int(sum(int(c.item()) for c in allc[: dist.get_rank()]))


def broadcast_unique_id(make_id=None) -> bytes | None:
    """NCCL unique id created on rank 0 (hsk_get_unique_id) and broadcast to every rank."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return None
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.zeros(capi.NCCL_ID_BYTES, dtype=torch.uint8, device=dev)
    if dist.get_rank() == 0:
        raw = (make_id or capi.get_unique_id)()
        t.copy_(torch.frombuffer(bytearray(raw), dtype=torch.uint8))
    dist.broadcast(t, 0)
    return bytes(t.cpu().numpy().tobytes())


def allreduce_histogram_host(hist: np.ndarray) -> np.ndarray:
    """host-side sum of per-rank histograms (alternative to hsk_allreduce_histogram)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return hist
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.from_numpy(hist.astype(np.int64)).to(dev)
    dist.all_reduce(t)
    return t.cpu().numpy().astype(np.uint64)


def create_context(k: int, m: int, lower: int, upper: int, ext: int = 0, **kw) -> capi.Context:
    """Context for this process's rank/GPU under torch.distributed (or single rank)."""
    import os
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size() > 1:
        local = int(os.environ.get("LOCAL_RANK", dist.get_rank()))
        return capi.Context(k, m, lower, upper, ext, device=local, rank=dist.get_rank(), nranks=dist.get_world_size(),
                            nccl_id=broadcast_unique_id(), **kw)
    return capi.Context(k, m, lower, upper, ext, **kw)
