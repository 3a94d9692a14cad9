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


def check_union(gathered, exp, hist, k, ext, world, what):
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
        if hist is not None:
            assert np.array_equal(hist, exp.hist), "all-reduced histogram"
        print(f"PASS {what} world={world} kept={got.n}", flush=True)
        return True
    except AssertionError as e:
        print(f"FAIL {what}: {e}", flush=True)
        return False


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    ok = True
    # (k, m, ext, read length, exchange mode, bins per rank: 0 = sized from the input -> small bins, supermer
    #  de-duplication across the sources of a bin; lower, error rate, coverage, HSK_TARGET_BIN, batch_kmers)
    cases = [(31, 17, 0, 150, "p2p", 64, 2, 0.01, 8.0, None, 0), (55, 23, 0, 2000, "p2p", 64, 2, 0.01, 8.0, None, 0),
             (31, 17, 1, 1000, "p2p", 64, 2, 0.01, 8.0, None, 0), (31, 17, 0, 3000, "p2p", 0, 2, 0.01, 30.0, None, 0),
             (31, 17, 0, 1000, "nccl", 0, 2, 0.01, 30.0, None, 0), (55, 23, 1, 400, "nccl", 64, 2, 0.01, 8.0, None, 0),
             # LOWER >= 5, bins sized from the input and large enough to keep more k-mers than a CTA sorts itself: the
             # staging area + big gather on several ranks (its cursor once shared words with the bin sizing)
             (31, 17, 0, 2000, "p2p", 0, 5, 0.001, 12.0, "30000", 0), (31, 17, 1, 2000, "p2p", 0, 5, 0.001, 12.0, "16000", 0),
             # every bin overflows its table: HBM path with one segment per source rank and bin, several batches
             (31, 17, 0, 2000, "p2p", 4, 2, 0.01, 12.0, None, 150_000), (31, 17, 1, 1000, "nccl", 4, 2, 0.01, 12.0, None, 150_000)]
    for (k, m, ext, read_len, mode, bpr, lower, err, cov, target, batch) in cases:
        os.environ["HSK_EXCHANGE"] = mode   # read when the context is created
        if target:
            os.environ["HSK_TARGET_BIN"] = target
        else:
            os.environ.pop("HSK_TARGET_BIN", None)
        rs = synth.sample_fixed(300_000, cov, read_len, err, seed=17 + k + ext)
        first = hd.partition_reads(rs.readlens, world)
        packed, lens, base = hd.shard(rs.packed, rs.readlens, first, rank)
        assert hd.readid_base(len(lens)) == base
        ctx = hd.create_context(k, m, lower, 50, ext, buckets_per_rank=bpr, batch_kmers=batch)
        r = ctx.count(packed, lens, readid_base=base)
        hist = ctx.allreduce_histogram()
        gathered = [None] * world
        dist.all_gather_object(gathered, {kk: r[kk] for kk in ("words", "cnt", "occ_off", "pos", "rid") if kk in r})
        st = r["stats"]
        print(f"[rank {rank}] k={k} ext={ext} L={lower} exchange={mode}: local k-mers {st['n_kmers_local']} owned {st['n_kmers_owned']} kept {r['n_kept']} "
              f"overflow bins {st['n_overflow_bins']} batches {st['n_batches']} sent {st['bytes_sent']} B recv {st['bytes_received']} B "
              f"exchange {st['ms_exchange']:.3f} ms", flush=True)
        if rank == 0:
            exp = po.kmer_count(rs.packed, rs.readlens, k, m, lower, 50, ext, via_supermers=False)
            ok = check_union(gathered, exp, hist, k, ext, world, f"k={k} ext={ext} L={lower} exchange={mode} bins/rank={bpr} target={target} batch={batch}") and ok
        ctx.close()
        dist.barrier()
    os.environ.pop("HSK_TARGET_BIN", None)
    os.environ["HSK_EXCHANGE"] = "p2p"

    # ---- the C++ API as a multi-rank job: hysortk::kmer_count(const DnaBuffer&, MPI_Comm) on every rank, the ranks of the
    #      bundled MPI stand-in (the NCCL id travels by MPI_Bcast, the ReadId base by MPI_Exscan: cxx/hysortk.cpp)
    from hysortk_b200 import cxxapi  # noqa: E402
    box = [os.urandom(6).hex()]
    dist.broadcast_object_list(box, src=0)
    os.environ.update(HSK_MPI_SIZE=str(world), HSK_MPI_RANK=str(rank), HSK_MPI_SESSION="mgpu" + box[0])
    for (k, m, ext) in [(31, 17, 0), (31, 17, 1), (55, 23, 0)]:
        rs = synth.sample_fixed(400_000, 15.0, 2500, 0.01, seed=3 + k + ext)
        first = hd.partition_reads(rs.readlens, world)
        packed, lens, base = hd.shard(rs.packed, rs.readlens, first, rank)
        a = cxxapi.kmer_count(packed, lens, k, m, 2, 50, ext)
        gathered = [None] * world
        dist.all_gather_object(gathered, {kk: a[kk] for kk in ("words", "cnt", "occ_off", "pos", "rid") if kk in a})
        if rank == 0:
            exp = po.kmer_count(rs.packed, rs.readlens, k, m, 2, 50, ext, via_supermers=False)
            ok = check_union(gathered, exp, None, k, ext, world, f"C++ API k={k} ext={ext}") and ok
        cxxapi.release(k, m, 2, 50, ext)
        dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MGPU_CHECK_OK" if ok else "MGPU_CHECK_FAILED", flush=True)
        sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
