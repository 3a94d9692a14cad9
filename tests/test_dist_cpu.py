"""world_size-2 gloo tests (CPU) of the multi-rank host logic in hysortk_b200/dist.py."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from hysortk_b200 import dist as hd
from hysortk_b200 import synth


def test_partition_matches_reference_rule():
    lens = np.array([100, 100, 100, 100, 100, 100, 100, 100], dtype=np.uint64)
    assert hd.partition_reads(lens, 2).tolist() == [0, 3, 8]     # "next read would reach the average" rule
    assert hd.partition_reads(lens, 1).tolist() == [0, 8]
    lens = np.array([1000, 10, 10, 10, 10], dtype=np.uint64)
    f = hd.partition_reads(lens, 3)
    assert f[0] == 0 and f[-1] == 5 and np.all(np.diff(f) >= 0)
    # shards tile the buffer exactly
    rs = synth.sample_mixed(5000, 40, [30, 31, 97, 150, 263], 0.0, seed=4)
    f = hd.partition_reads(rs.readlens, 3)
    parts = [hd.shard(rs.packed, rs.readlens, f, r) for r in range(3)]
    assert np.array_equal(np.concatenate([p[0] for p in parts]), rs.packed)
    assert np.array_equal(np.concatenate([p[1] for p in parts]), rs.readlens)
    assert [p[2] for p in parts] == f[:3].tolist()


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        base = hd.readid_base(10 + 5 * rank)
        uid = hd.broadcast_unique_id(make_id=lambda: bytes(range(128)))
        h = hd.allreduce_histogram_host(np.arange(8, dtype=np.uint64) * (rank + 1))
        q.put((rank, base, uid, h.tolist()))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_host_logic():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == 0 and res[1][1] == 10                  # exclusive scan of the read counts
    assert res[0][2] == res[1][2] == bytes(range(128))         # same id on every rank
    assert res[0][3] == res[1][3] == [3 * i for i in range(8)]  # summed histogram
