"""Multi-GPU parity check, run under torchrun (one rank per GPU):
   python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/mgpu_check.py
Every rank counts its contiguous share of the reads; the union over ranks must equal the oracle's
result on the whole read set (partition independence), the per-rank results must be
disjoint, and the all-reduced histogram must equal the oracle's."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hysortk_b200 import dist as hd  # noqa: E402
from hysortk_b200 import synth  # noqa: E402
from oracle import pyoracle as po  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    ok = True
    # (k, m, ext, read length, exchange mode, bins per rank: 0 = sized from the input -> small bins, supermer
    # de-duplication across the sources of a bin)
    for (k, m, ext, read_len, mode, bpr) in [(31, 17, 0, 150, "p2p", 64), (55, 23, 0, 2000, "p2p", 64), (31, 17, 1, 1000, "p2p", 64),
                                             (31, 17, 0, 3000, "p2p", 0), (31, 17, 0, 1000, "nccl", 0), (55, 23, 1, 400, "nccl", 64)]:
        os.environ["HSK_EXCHANGE"] = mode   # read when the context is created
        rs = synth.sample_fixed(300_000, 8.0 if bpr else 30.0, read_len, 0.01, seed=17 + k + ext)
        first = hd.partition_reads(rs.readlens, world)
        packed, lens, base = hd.shard(rs.packed, rs.readlens, first, rank)
        assert hd.readid_base(len(lens)) == base
        ctx = hd.create_context(k, m, 2, 50, ext, buckets_per_rank=bpr)
        r = ctx.count(packed, lens, readid_base=base)
        hist = ctx.allreduce_histogram()
        gathered = [None] * world
        dist.all_gather_object(gathered, {kk: r[kk] for kk in ("words", "cnt", "occ_off", "pos", "rid") if kk in r})
        st = r["stats"]
        print(f"[rank {rank}] k={k} ext={ext} exchange={mode}: local k-mers {st['n_kmers_local']} owned {st['n_kmers_owned']} kept {r['n_kept']} "
              f"sent {st['bytes_sent']} B recv {st['bytes_received']} B exchange {st['ms_exchange']:.3f} ms", flush=True)
        if rank == 0:
            exp = po.kmer_count(rs.packed, rs.readlens, k, m, 2, 50, ext, via_supermers=False)
            words = np.concatenate([g["words"] for g in gathered])
            cnt = np.concatenate([g["cnt"] for g in gathered])
            if ext:
                occ_off = [np.zeros(1, dtype=np.uint64)]
                shift = 0
                for g in gathered:
                    occ_off.append(g["occ_off"][1:] + np.uint64(shift))
                    shift += int(g["occ_off"][-1])
                got = po.canonicalize(k, words, cnt, np.concatenate(occ_off), np.concatenate([g["pos"] for g in gathered]),
                                      np.concatenate([g["rid"] for g in gathered]))
            else:
                got = po.canonicalize(k, words, cnt)
            try:
                po.assert_equal(got, exp, f"{world}-GPU union vs oracle")
                assert len(np.unique(words, axis=0)) == len(words), "per-rank results overlap"
                assert np.array_equal(hist, exp.hist), "all-reduced histogram"
                print(f"PASS k={k} ext={ext} exchange={mode} world={world} kept={got.n}", flush=True)
            except AssertionError as e:
                ok = False
                print(f"FAIL k={k} ext={ext}: {e}", flush=True)
        ctx.close()
        dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MGPU_CHECK_OK" if ok else "MGPU_CHECK_FAILED", flush=True)
        sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
