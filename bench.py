#!/usr/bin/env python
"""Benchmark of the kmer_count hot path (BASELINE.json metric: k-mers counted / second, whole job).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W

A "step" is one full kmer_count over the synthetic read set of the workload; the default is BASELINE.json configs[1]
(K=31 M=17 L=2 U=50, 30x reads of a 5 Mbp uniform genome, ~150 Mbp, 1 % substitutions) per GPU.  With N GPUs the genome is
N times larger and every rank holds its own share of the reads (weak scaling); the supermer all-to-all is fused into the
count kernel (bins read the peers' supermer streams in place over NVLink; HSK_EXCHANGE=nccl selects the grouped
ncclSend/ncclRecv baseline).  Other workloads: `c3_share` / `c4_share` / `c5_share` = the per-GPU share of BASELINE.json
configs[2..4] (3.75 Gbp K=31; 3.75 Gbp K=55 M=23; 1.25 Gbp K=31 EXT=1), reads generated on the GPU; `c3_30Gbp` =
configs[2] as a whole (30 Gbp split over the ranks: strong scaling).

  parity_check  before anything is timed: a ~20 Mbp read set with the workload's K/M/L/U/EXT is counted by all ranks and the
         union of the per-rank results is compared with the oracle (oracle/oracle.c) on rank 0, bit for bit.
  value  device path: reads resident in HBM when the timed region starts, result left in HBM
         (hsk_count_device), timed with CUDA events on the launching stream, max over ranks.
  e2e    the same metric through hysortk::kmer_count(const DnaBuffer&, MPI_Comm) itself — the call an ELBA-style caller
         makes (include/hysortk.hpp; C entry points hysortk_b200/cxx/bench_api.cpp): pageable DnaBuffer in, staged H2D
         copies, count, D2H of the result, std::vector<KmerListEntryS> built on the host.  Wall clock of the call on
         every rank (ranks of the bundled multi-process MPI stand-in, one per GPU), max over ranks.
         e2e_pinned_soa: the C-ABI call hsk_count on page-locked input, result left as page-locked arrays (no KmerListS).
         e2e_ceiling: the same bytes as bare concurrent cudaMemcpyAsync H2D + D2H on page-locked memory, all ranks at once.
  roofline   dominant kernel = k_bin_count (stages 4+5 fused on chip).  frac = bytes the kernel has to move (supermers in,
             kept entries out) / its launch time / measured HBM copy bandwidth: it is NOT HBM-bound (issue slots +
             shared-memory wavefronts limit it, ncu summaries under profiles/), so the fraction is small by design;
             survey_8d_* = the SURVEY.md 8(d) figure for the HBM-resident expand + 8-bit LSD sort + count it replaces.
  cpu_baseline  the UNMODIFIED reference (oracle/_ref, built from /root/reference with the bundled MPI stand-in) on this
             box's host cores, as an MPI job of R ranks x T OpenMP threads (the best of a short calibration).

`--impl reference` times that reference build as its own arm (rank 0 only starts it; it uses all host cores).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
import uuid

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K, M, LOWER, UPPER, EXT = 31, 17, 2, 50, 0
METRIC = "kmers_counted_per_sec"
UNIT = "kmers/s"

WORKLOADS = {
    # genome_len: bases of genome PER GPU (weak scaling: N ranks -> N times the genome, every rank its own reads)
    "c2_150Mbp_10kbp": dict(genome_len=5_000_000, coverage=30.0, read_len=10_000, err=0.01, gen="host"),
    "c2_150Mbp_150bp": dict(genome_len=5_000_000, coverage=30.0, read_len=150, err=0.01, gen="host"),
    "c1_100Mbp_150bp": dict(genome_len=3_340_000, coverage=30.0, read_len=150, err=0.01, gen="host"),
    "tiny": dict(genome_len=200_000, coverage=10.0, read_len=1000, err=0.01, gen="host"),
    # per-GPU shares of BASELINE.json configs[2..4] at 8 GPUs
    "c3_share": dict(genome_len=375_000_000, coverage=10.0, read_len=10_000, err=0.01, gen="gpu"),
    "c4_share": dict(genome_len=125_000_000, coverage=30.0, read_len=10_000, err=0.01, gen="gpu", k=55, m=23),
    "c5_share": dict(genome_len=62_500_000, coverage=20.0, read_len=10_000, err=0.01, gen="gpu", ext=1),
    # configs[2] as a whole: 10x reads of a 3 Gbp genome, split over the ranks (strong scaling)
    "c3_30Gbp": dict(genome_len=3_000_000_000, coverage=10.0, read_len=10_000, err=0.01, gen="gpu", strong=True),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2_150Mbp_10kbp", choices=sorted(WORKLOADS))
    ap.add_argument("--batch-kmers", type=int, default=0)
    ap.add_argument("--buckets-per-rank", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="device path only (huge inputs: no host copy of the reads)")
    ap.add_argument("--cpu-sample-mbp", type=float, default=60.0)
    ap.add_argument("--ref-budget-s", type=float, default=150.0, help="CPU seconds the reference arm may spend in total")
    # the compile-time parameters of the reference (defaults: the workload's, else BASELINE.json configs[1])
    ap.add_argument("--k", type=int, default=None)
    ap.add_argument("--m", type=int, default=None)
    ap.add_argument("--lower", type=int, default=LOWER)
    ap.add_argument("--upper", type=int, default=UPPER)
    ap.add_argument("--ext", type=int, default=None)
    a = ap.parse_args()
    w = WORKLOADS[a.workload]
    globals().update(K=a.k if a.k is not None else w.get("k", K), M=a.m if a.m is not None else w.get("m", M), LOWER=a.lower,
                     UPPER=a.upper, EXT=a.ext if a.ext is not None else w.get("ext", EXT))
    return a


def workload_config(name: str, nranks: int, seed: int = 42) -> dict:
    """The keys both arms print under `config` (what the workload IS; nothing about how it was run)."""
    p = WORKLOADS[name]
    strong = bool(p.get("strong"))
    G_total = p["genome_len"] if strong else p["genome_len"] * nranks
    reads_total = int(G_total * p["coverage"] / p["read_len"])
    reads_per_gpu = reads_total // nranks
    return {"workload": name, "k": K, "m": M, "lower": LOWER, "upper": UPPER, "ext": EXT, "genome_len_total": G_total,
            "coverage": p["coverage"], "read_len": p["read_len"], "err": p["err"], "seed": seed,
            "reads_per_gpu": reads_per_gpu, "kmers_per_gpu": reads_per_gpu * max(p["read_len"] - K + 1, 0),
            "scaling": "strong" if strong else "weak"}


def make_shard_host(genome_len_total: int, nreads: int, read_len: int, err: float, rank: int, seed: int = 42):
    """`nreads` reads of this rank, sampled from the genome all ranks share (numpy; rank-specific stream)."""
    from hysortk_b200 import synth
    genome = synth.make_genome(genome_len_total, seed)
    rng = np.random.Generator(np.random.Philox(seed + 1000 + rank))
    L = read_len
    nb = (L + 3) // 4
    packed = np.empty(nreads * nb, dtype=np.uint8)
    ar = np.arange(L, dtype=np.int64)
    chunk = max(1, (1 << 24) // L)
    for s in range(0, nreads, chunk):
        e = min(nreads, s + chunk)
        starts = rng.integers(0, genome_len_total - L + 1, size=e - s, dtype=np.int64)
        reads = genome[starts[:, None] + ar[None, :]]
        reads = synth._mutate_and_flip(reads, err, rng)
        packed[s * nb:e * nb] = synth.pack_codes_matrix(reads).reshape(-1)
    return synth.ReadSet(packed, np.full(nreads, L, dtype=np.uint64))


def make_shard_gpu(genome_len_total: int, nreads: int, read_len: int, err: float, rank: int, dev, seed: int = 42):
    """The same distribution generated on the device (large inputs): returns (d_packed, d_off, d_len, nbytes)."""
    import torch
    g = torch.Generator(device=dev).manual_seed(seed)
    G = genome_len_total
    genome = torch.empty(G, dtype=torch.uint8, device=dev)
    for s in range(0, G, 1 << 28):   # the genome is the same on every rank
        e = min(G, s + (1 << 28))
        genome[s:e] = torch.randint(0, 4, (e - s,), dtype=torch.uint8, device=dev, generator=g)
    g = torch.Generator(device=dev).manual_seed(seed + 1000 + rank)
    nb = (read_len + 3) // 4
    packed = torch.zeros(nreads * nb + 64, dtype=torch.uint8, device=dev)
    ar = torch.arange(read_len, device=dev)
    chunk = max(1, (1 << 26) // read_len)
    for s in range(0, nreads, chunk):
        e = min(nreads, s + chunk)
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
        del r, mut, flip, starts
    del genome
    torch.cuda.empty_cache()
    off = torch.arange(nreads + 1, dtype=torch.int64, device=dev) * nb
    lens = torch.full((nreads,), read_len, dtype=torch.int32, device=dev)
    return packed, off, lens, nreads * nb


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.device), "-lms", "50"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


# ------------------------------------------------------------------------------------- the reference on the host cores

def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def rank_thread_shapes(cores: int):
    """MPI ranks x OpenMP threads the reference is tried with (its README runs one rank per NUMA domain; with one rank
    its expansion of the received supermers is serial, kmerops.cpp:997-1004)."""
    shapes = [(1, cores)]
    for r in (2, 4, 8):
        if cores // r >= 4:
            shapes.append((r, cores // r))
    return shapes


def run_reference(sample, steps: int, warmup: int, shapes=None):
    """Times the unmodified reference's kmer_count (oracle/_ref) on `sample` (a ReadSet) as an MPI job on this box:
    a short calibration picks ranks x threads, then warmup + steps runs of the whole sample."""
    from oracle import pyoracle as po
    if not po.ref_available(K, M, LOWER, UPPER, EXT):
        return None
    cores = host_cores()
    shapes = shapes or rank_thread_shapes(cores)
    nk = sample.num_kmers(K)
    if len(shapes) > 1:
        L = int(sample.readlens[0])
        ncal = max(shapes[-1][0] * 2, min(sample.nreads, int(8e6 // L)))   # ~8 Mbp, at least two reads per rank
        nb = (L + 3) // 4
        cal_p, cal_l = sample.packed[: ncal * nb], sample.readlens[:ncal]
        best = None
        for (r, t) in shapes:
            try:
                _, secs = po.ref_kmer_count_ranks(cal_p, cal_l, K, M, LOWER, UPPER, EXT, nranks=r, threads_per_rank=t,
                                                  want_result=False, repeats=2)
            except Exception as e:   # a shape the reference cannot run (too few reads per rank, ...)
                print(f"[bench] reference shape {r}x{t} failed: {e}", file=sys.stderr)
                continue
            if best is None or min(secs) < best[0]:
                best = (min(secs), r, t)
        if best is None:
            return None
        _, r, t = best
    else:
        r, t = shapes[0]
    _, secs = po.ref_kmer_count_ranks(sample.packed, sample.readlens, K, M, LOWER, UPPER, EXT, nranks=r, threads_per_rank=t,
                                      want_result=False, repeats=warmup + steps)
    sec = float(np.mean(secs[warmup:]))
    return dict(value=nk / sec, seconds_per_step=sec, kmers=nk, cores=r * t, ranks=r, threads=t,
                sample=f"{sample.nreads} reads ({sample.nbases / 1e6:.1f} Mbp, {nk} k-mers) at the workload's coverage, "
                       f"{r} MPI rank(s) x {t} OpenMP threads (best of {len(shapes)} shapes on a short calibration), RADULS")


def reference_sample(cfg: dict, nranks: int, max_kmers: float):
    """Reads for the reference arm: the whole workload of all ranks when it fits the budget, else as many reads at the
    same coverage over a proportionally smaller genome (the k-mer spectrum, which decides the sort / count cost, is
    kept)."""
    L = cfg["read_len"]
    total_reads = cfg["reads_per_gpu"] * nranks
    per_read = max(L - K + 1, 1)
    n = int(min(total_reads, max(nranks * 2, max_kmers // per_read)))
    G = cfg["genome_len_total"] if n == total_reads else max(4 * L, int(n * L / cfg["coverage"]))
    return make_shard_host(G, n, L, cfg["err"], 0, cfg["seed"]), n == total_reads


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nranks = max(1, args.gpus)
    cfg = workload_config(args.workload, nranks)
    steps, warmup = args.steps, args.warmup
    # every step = the reference on the workload of all N ranks, or on a bounded sample of it when warmup + steps runs
    # of the whole would exceed the CPU budget (the reference counts ~40 M k-mers/s on a node of this class)
    budget_kmers = args.ref_budget_s * 40e6 / max(1, steps + warmup)
    sample, whole = reference_sample(cfg, nranks, budget_kmers)
    r = run_reference(sample, steps, warmup)
    if r is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built (needs /root/reference; run oracle/build_ref.sh)"}))
        return
    r["sample"] = ("the whole workload: " if whole else "a bounded sample of the workload: ") + r["sample"]
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": r["seconds_per_step"] * 1e3, "higher_is_better": True, "scaling": cfg["scaling"],
            "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "reference", "sample": r["sample"],
                             "ranks": r["ranks"], "threads_per_rank": r["threads"]},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------ our arm

def parity_check(dist, hd, capi, rank: int, world: int, dev) -> dict:
    """All ranks count a small read set with this run's parameters; rank 0 compares the union with the oracle."""
    from hysortk_b200 import synth
    from oracle import pyoracle as po
    rs = synth.sample_fixed(1_000_000, 20.0, 2000, 0.01, seed=97 + K + EXT)
    first = hd.partition_reads(rs.readlens, world)
    packed, lens, base = hd.shard(rs.packed, rs.readlens, first, rank)
    ctx = hd.create_context(K, M, LOWER, UPPER, EXT) if world > 1 else capi.Context(K, M, LOWER, UPPER, EXT, device=dev.index)
    r = ctx.count(packed, lens, readid_base=base)
    hist = ctx.allreduce_histogram()
    part = {kk: r[kk] for kk in ("words", "cnt", "occ_off", "pos", "rid") if kk in r}
    gathered = [None] * world
    if world > 1:
        dist.gather_object(part, gathered if rank == 0 else None, dst=0)
    else:
        gathered = [part]
    ctx.close()
    res = {"status": "ok", "against": "oracle/oracle.c (C restatement of the reference, pinned to the reference's own output)",
           "reads_mbp": rs.nbases / 1e6, "ranks": world}
    if rank == 0:
        exp = po.kmer_count(rs.packed, rs.readlens, K, M, LOWER, UPPER, EXT, via_supermers=False)
        words = np.concatenate([g["words"] for g in gathered])
        cnt = np.concatenate([g["cnt"] for g in gathered])
        if EXT:
            offs, shift = [np.zeros(1, dtype=np.uint64)], 0
            for g in gathered:
                offs.append(g["occ_off"][1:] + np.uint64(shift))
                shift += int(g["occ_off"][-1])
            got = po.canonicalize(K, words, cnt, np.concatenate(offs), np.concatenate([g["pos"] for g in gathered]),
                                  np.concatenate([g["rid"] for g in gathered]))
        else:
            got = po.canonicalize(K, words, cnt)
        try:
            po.assert_equal(got, exp, f"{world}-rank union vs oracle")
            assert len(np.unique(words, axis=0)) == len(words), "per-rank results overlap"
            assert np.array_equal(hist, exp.hist), "all-reduced histogram"
            res["kept"] = int(got.n)
        except AssertionError as e:
            res = {"status": "FAILED", "error": str(e)[:300], "ranks": world}
    if world > 1:
        box = [res]
        dist.broadcast_object_list(box, src=0)
        res = box[0]
    return res


def copy_ceiling(torch, dist, dev, world: int, h2d_bytes: int, d2h_bytes: int, reps: int = 20) -> float:
    """Seconds per step of the bare copies of a step: H2D + D2H of page-locked buffers, concurrently, on all ranks."""
    hin = torch.empty(max(h2d_bytes, 1), dtype=torch.uint8).pin_memory()
    hout = torch.empty(max(d2h_bytes, 1), dtype=torch.uint8).pin_memory()
    din = torch.empty(max(h2d_bytes, 1), dtype=torch.uint8, device=dev)
    dout = torch.empty(max(d2h_bytes, 1), dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def once():
        with torch.cuda.stream(s1):
            din.copy_(hin, non_blocking=True)
        with torch.cuda.stream(s2):
            hout.copy_(dout, non_blocking=True)

    for _ in range(3):
        once()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    return dt


def main_ours(args):
    import torch
    import torch.distributed as dist
    from hysortk_b200 import capi, cxxapi, synth
    from hysortk_b200 import dist as hd

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.gpus > 1 and world == 1:
        raise SystemExit("for --gpus N > 1 launch with torch.distributed.run (one rank per GPU)")
    if world > 1:
        # one rank per GPU: keep the rank's host thread (and so its page-locked buffers, first touch) on the CPUs next to
        # its GPU; without it half of the ranks stage their H2D / D2H copies through the other socket
        try:
            import pynvml
            pynvml.nvmlInit()
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
        except Exception as e:   # not fatal: only the host-buffer (e2e) path is affected
            print(f"[bench] no CPU affinity for rank {rank}: {e}", file=sys.stderr)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    nccl_id = None
    session = uuid.uuid4().hex[:12]
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        box = [session]
        dist.broadcast_object_list(box, src=0)
        session = box[0]
        nccl_id = hd.broadcast_unique_id()
    host_threads = max(1, host_cores() // world)

    cfg = workload_config(args.workload, world)
    wl = WORKLOADS[args.workload]

    # ---- parity first: nothing is timed on a build whose multi-rank result differs from the oracle's ----------------
    parity = {"status": "skipped"}
    if not args.no_parity:
        parity = parity_check(dist, hd, capi, rank, world, dev)
        if parity["status"] != "ok":
            if rank == 0:
                print(json.dumps({"metric": METRIC, "value": 0.0, "unit": UNIT, "n_gpus": world, "parity_check": parity,
                                  "error": "parity check failed: nothing was timed"}))
            raise SystemExit(3)

    # ---- the reads of this rank ----------------------------------------------------------------------------------------
    nreads, L = cfg["reads_per_gpu"], cfg["read_len"]
    nk_local = cfg["kmers_per_gpu"]
    want_host = not args.no_e2e
    if wl["gen"] == "host":
        rs = make_shard_host(cfg["genome_len_total"], nreads, L, cfg["err"], rank, cfg["seed"])
        nbytes = rs.packed.nbytes
        d_packed = torch.zeros(((nbytes + 15) // 16) * 16 + 64, dtype=torch.uint8, device=dev)
        d_packed[:nbytes].copy_(torch.from_numpy(rs.packed))
        d_off = torch.from_numpy(rs.byte_offsets().view(np.int64)).to(dev)
        d_len = torch.from_numpy(rs.readlens.astype(np.uint32).view(np.int32)).to(dev)
        h_packed_np, h_lens_np = rs.packed, rs.readlens
    else:
        d_packed, d_off, d_len, nbytes = make_shard_gpu(cfg["genome_len_total"], nreads, L, cfg["err"], rank, dev, cfg["seed"])
        h_packed_np = d_packed[:nbytes].cpu().numpy() if want_host else None
        h_lens_np = np.full(nreads, L, dtype=np.uint64)
    stream = torch.cuda.current_stream()
    ctx = capi.Context(K, M, LOWER, UPPER, EXT, device=local_rank, rank=rank, nranks=world, nccl_id=nccl_id,
                       buckets_per_rank=args.buckets_per_rank, batch_kmers=args.batch_kmers, stream=stream.cuda_stream)
    readid_base = rank * nreads

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        return ctx.count_device(d_packed.data_ptr(), nbytes, d_off.data_ptr(), d_len.data_ptr(), nreads, readid_base)

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- device path --------------------------------------------------------------------------------------------------
    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stats_acc = []
    e0.record(stream)
    for _ in range(args.steps):
        r = step_device()
        stats_acc.append(r.stats.as_dict())
    e1.record(stream)
    barrier()
    ms_dev = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None
    n_kept, n_occ = int(r.n_kept), int(r.n_occ)
    if stats_acc[-1]["n_kmers_local"] != nk_local:
        raise SystemExit(f"k-mer total mismatch: engine {stats_acc[-1]['n_kmers_local']} vs {nk_local}")
    total_kmers = sum_over_ranks(float(nk_local))
    value = total_kmers * args.steps / (ms_dev * 1e-3)
    st = {k: float(np.mean([s[k] for s in stats_acc])) for k in stats_acc[0]}

    # ---- end to end ---------------------------------------------------------------------------------------------------
    h2d = nbytes + nreads * 8
    d2h = n_kept * (8 * ctx.nwords + 4 + (8 if EXT else 0)) + n_occ * 8 * (1 if EXT else 0) + (UPPER + 1) * 8 + 16
    e2e = None
    e2e_pinned = None
    ceiling = None
    e2e_steps = max(1, min(args.steps, 50))
    if want_host:
        # (1) C ABI on page-locked buffers, result left as page-locked arrays
        h_packed = torch.from_numpy(h_packed_np).pin_memory()
        h_lens = torch.from_numpy(h_lens_np.view(np.int64)).pin_memory()

        def step_pinned():
            rr = capi.Result()
            capi._check(ctx.lib.hsk_count(ctx.handle, h_packed.data_ptr(), nbytes, h_lens.data_ptr(), nreads, readid_base,
                                          capi.C.byref(rr)))
            return rr

        for _ in range(min(args.warmup, 3)):
            step_pinned()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step_pinned()
        torch.cuda.synchronize()
        ms_pin = max_over_ranks((time.perf_counter() - t0) * 1e3)
        e2e_pinned = {"value": total_kmers * e2e_steps / (ms_pin * 1e-3), "unit": UNIT, "ms_per_step": ms_pin / e2e_steps,
                      "what": "hsk_count (C ABI) on page-locked input, result as page-locked arrays: no KmerListS"}
        del h_packed, h_lens
        ceiling_s = copy_ceiling(torch, dist, dev, world, h2d, d2h)
        ceiling = {"value": total_kmers / ceiling_s, "unit": UNIT, "ms_per_step": ceiling_s * 1e3,
                   "what": "bare cudaMemcpyAsync of the step's H2D + D2H bytes on page-locked buffers, both directions at once, all ranks"}
    ctx.close()   # the C++ API owns its own engine context: one at a time on the GPU
    if want_host:
        # (2) THE end-to-end number: hysortk::kmer_count on a pageable DnaBuffer, KmerListS out
        if world > 1:
            os.environ.update(HSK_MPI_SIZE=str(world), HSK_MPI_RANK=str(rank), HSK_MPI_SESSION="bench" + session)
        barrier()
        res = cxxapi.bench(h_packed_np, h_lens_np, K, M, LOWER, UPPER, EXT, warmup=min(args.warmup, 3), steps=e2e_steps,
                           threads=host_threads)
        ms_api = max_over_ranks(float(res["seconds"].sum()) * 1e3)
        if res["n_kept"] != n_kept:
            raise SystemExit(f"kmer_count through the C++ API kept {res['n_kept']} entries, the device path {n_kept}")
        cxxapi.release(K, M, LOWER, UPPER, EXT)
        e2e = {"value": total_kmers * e2e_steps / (ms_api * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": ms_api / e2e_steps, "steps": e2e_steps, "host_threads_per_rank": host_threads,
               "what": "hysortk::kmer_count(const DnaBuffer&, MPI_Comm): pageable DnaBuffer in, std::vector<KmerListEntryS> out; "
                       "wall clock of the call, max over ranks"}
    else:
        e2e = {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "what": "not measured (--no-e2e)"}

    # ---- roofline of the dominant kernel ----------------------------------------------------------------------------
    rec = 8 * ctx.nwords + (8 if EXT else 0)
    n_owned = st["n_kmers_owned"]
    nw = ctx.nwords
    npass = (2 * (K - 32 * (nw - 1)) + 7) // 8 + 8 * (nw - 1)
    peak, peak_kind = measured_hbm_peak()
    # SURVEY.md 8(d) algorithmic bytes of the stages the fused on-chip kernel covers: expand (supermer bytes + N*rec
    # written), LSD sort N*rec*(1+2P), count (N*rec read + D*(W+4) written)
    survey_bytes = st["supermer_bytes"] + n_owned * rec + n_owned * rec * (1 + 2 * npass) + n_owned * rec + n_kept * (8 * nw + 4)
    need_bytes = st["supermer_bytes"] + n_kept * (8 * nw + 4) + (n_occ * 8 if EXT else 0)
    ms_bins = st["ms_bins"]
    achieved = need_bytes / (ms_bins * 1e-3) / 1e9 if ms_bins > 0 else 0.0
    traffic = issue_active = warp_inst = None
    try:   # per-launch dram bytes / issue-slot utilisation / executed warp instructions of the kernel from the committed
           # ncu --set full capture
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        traffic = tj.get(f"k_bin_count:{args.workload}:k{K}:ext{EXT}")
        issue_active = tj.get(f"k_bin_count:{args.workload}:k{K}:ext{EXT}:issue_active_pct")
        warp_inst = tj.get(f"k_bin_count:{args.workload}:k{K}:ext{EXT}:warp_instructions")
    except (OSError, ValueError):
        pass
    # the kernel's own ceiling: one warp instruction per scheduler and clock (4 schedulers per SM)
    issue = None
    if warp_inst and ms_bins > 0:
        props = torch.cuda.get_device_properties(dev)
        clock_mhz = (clocks or {}).get("sm_mhz") or (clocks or {}).get("sm_max_mhz") or 1965.0
        peak_issue = props.multi_processor_count * 4 * clock_mhz * 1e6
        issue = {"warp_instructions_per_launch": warp_inst, "achieved_warp_inst_per_s": warp_inst / (ms_bins * 1e-3),
                 "peak_warp_inst_per_s": peak_issue, "frac": warp_inst / (ms_bins * 1e-3) / peak_issue,
                 "note": "executed warp instructions of the ncu capture over this run's launch time, against SMs x 4 schedulers x SM clock"}
    roofline = {"bound": "issue+smem",
                "kernel": "k_bin_count (expand + hash-count + sort one bin per CTA in shared memory; supermers read in place "
                          "from local HBM or from the peers over NVLink)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_kind, "algorithmic_bytes_per_launch": need_bytes, "ms_per_launch": ms_bins, "launches_per_step": 1,
                "issue_active_pct_ncu": issue_active, "issue": issue,
                "note": "algorithmic bytes = what the kernel must move through HBM per launch: the supermer slots of the owned bins "
                        "in, the kept (k-mer, count[, occurrence]) entries out; the k-mer occurrences never leave the SM, so the "
                        "kernel is bound by instruction issue and shared-memory wavefronts, not by HBM (ncu: profiles/)",
                "survey_8d_bytes_per_launch": survey_bytes,
                "survey_8d_equivalent_gbs": survey_bytes / (ms_bins * 1e-3) / 1e9 if ms_bins > 0 else 0.0,
                "survey_8d_note": "SURVEY.md 8(d) bytes of the HBM-resident stages this kernel replaces (expand + 8-bit LSD sort + "
                                  "count); a rate above the HBM peak only says that no HBM-resident sort could be as fast",
                "hbm_path": {"overflow_bins": st["n_overflow_bins"], "ms_expand": st["ms_expand"], "ms_sort": st["ms_sort"],
                             "ms_count": st["ms_count"]},
                "stage_ms": {k: st[k] for k in ["ms_extract", "ms_exchange", "ms_bins", "ms_expand", "ms_sort", "ms_count", "ms_total"]}}
    exchange = None
    if world > 1:
        mode = "nccl send/recv" if os.environ.get("HSK_EXCHANGE") == "nccl" else "fused: bins read peer streams in place (P2P loads over NVLink)"
        t_x = st["ms_exchange"] if os.environ.get("HSK_EXCHANGE") == "nccl" else st["ms_bins"]
        exchange = {"mode": mode, "bytes_received_per_rank": st["bytes_received"], "ms_exchange_only": st["ms_exchange"],
                    "ms_window": t_x, "gbs_per_rank": st["bytes_received"] / (t_x * 1e-3) / 1e9 if t_x > 0 else 0.0,
                    "nvlink_peak_gbs_per_dir": 900.0,
                    "note": "fused mode: ms_exchange_only = all-gathers of the bin totals / IPC records + the barrier; the "
                            "supermers cross NVLink inside k_bin_count, so gbs_per_rank is bytes over that kernel's time"}
    # ---- CPU baseline (rank 0, N=1 only) ----------------------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample, _ = reference_sample(cfg, 1, args.cpu_sample_mbp * 1e6)
        rr = run_reference(sample, 2, 1)
        if rr is not None:
            cpu = {"value": rr["value"], "unit": UNIT, "cores": rr["cores"], "kind": "reference", "sample": rr["sample"],
                   "ranks": rr["ranks"], "threads_per_rank": rr["threads"]}
        else:
            from oracle import pyoracle as po
            n = max(1, int(5e6 // L)); nb = (L + 3) // 4
            t0 = time.time()
            c = po.kmer_count(sample.packed[: n * nb], sample.readlens[:n], K, M, LOWER, UPPER, EXT, ntasks=5)
            dt = time.time() - t0
            cpu = {"value": c.total_kmers / dt, "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": f"first {n} reads ({n * L / 1e6:.1f} Mbp) through oracle/oracle.c (scalar)"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None,
                "dtype": "u64", "data": "synthetic", "config": cfg,
                "details": {"kept_kmers_rank0": n_kept, "bins_per_rank": args.buckets_per_rank or "auto",
                            "overflow_bins": st["n_overflow_bins"],
                            "l2_policy": "buffers written every step (run list, supermers, results: hundreds of MB) exceed the 126 MB L2; no flush",
                            "reads_generated_on": wl["gen"]},
                "parity_check": parity, "clocks": clocks, "e2e": e2e, "e2e_pinned_soa": e2e_pinned, "e2e_ceiling": ceiling,
                "gpu_launches": int(round(st["n_launches"] * args.steps)),
                "roofline": roofline, "cpu_baseline": cpu}
        if exchange:
            line["exchange"] = exchange
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    # the driver parses ONE JSON line from stdout: route everything else (NCCL banners, library chatter) to stderr
    sys.stdout.flush()
    _real_stdout = os.dup(1)
    os.dup2(2, 1)
    _out = os.fdopen(_real_stdout, "w")
    _print = print

    def print(*args, **kw):  # noqa: A001 - the final JSON line goes to the real stdout
        kw.setdefault("file", _out)
        _print(*args, **kw)
        _out.flush()

    if a.impl == "reference":
        main_reference(a)
    else:
        main_ours(a)
