"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI
(include/hsk_capi.h), against the oracle, the golden fixtures of the unmodified reference, and
size-independent properties at larger sizes."""
import numpy as np
import pytest

from hysortk_b200 import capi, synth
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


def gpu_counts(ctx, rs_packed, readlens, readid_base=0):
    r = ctx.count(rs_packed, readlens, readid_base)
    c = po.canonicalize(ctx.k, r["words"], r["cnt"], r.get("occ_off"), r.get("pos"), r.get("rid"))
    return c, r


def slot_codes(sm, s):
    """2-bit codes of the bases of slot s (payload words, 16 bases per word from the top bits)."""
    w = sm["payload"][s].copy()
    w[-1] &= np.uint32(0xFFFFFF00)
    codes = np.stack([(w >> np.uint32(30 - 2 * j)) & np.uint32(3) for j in range(16)], axis=1).reshape(-1)
    return codes[: int(sm["len"][s])].astype(np.uint8)


def expand_supermers_cpu(sm, k):
    """numpy re-expansion of the supermer slots returned by hsk_debug_extract: (bin, canonical k-mer words)
    for every k-mer of every slot."""
    nw = 1 if k <= 32 else (2 if k <= 64 else 3)
    lens = sm["len"]
    bucket_of = np.repeat(np.arange(sm["n_buckets"]), sm["bucket_count"].astype(np.int64))
    out_b, out_w = [], []
    for s in range(len(lens)):
        codes = slot_codes(sm, s)
        for i in range(int(lens[s]) - k + 1):
            f = codes[i:i + k]
            r = (3 - f[::-1]).astype(np.uint8)
            c = f if tuple(f) <= tuple(r) else r
            ww = [0] * nw
            for j, b in enumerate(c):
                ww[j // 32] |= int(b) << (2 * (31 - j % 32))
            out_b.append(bucket_of[s]); out_w.append(ww)
    return np.array(out_b), np.array(out_w, dtype=np.uint64).reshape(-1, nw)


@pytest.mark.parametrize("k,m,ext", [(31, 17, 0), (55, 23, 0), (31, 17, 1), (21, 11, 0), (77, 29, 0)])
def test_extract_supermers_cover_all_kmers(k, m, ext):
    rs = synth.sample_mixed(5000, 60, [k - 1, k, k + 1, 97, 150, 263, 1000], 0.02, seed=k)
    with capi.Context(k, m, 1, 65535, ext, buckets_per_rank=16) as ctx:
        sm = ctx.debug_extract(rs.packed, rs.readlens, readid_base=5)
    assert int(sm["bucket_kmers"].sum()) == rs.num_kmers(k)
    assert int(sm["bucket_count"].sum()) == len(sm["len"])
    b, w = expand_supermers_cpu(sm, k)
    assert len(b) == rs.num_kmers(k)
    # per bucket k-mer totals as reported
    assert np.array_equal(np.bincount(b, minlength=sm["n_buckets"]).astype(np.uint64), sm["bucket_kmers"])
    # the multiset of canonical k-mers equals the direct definition
    exp = po.kmer_count(rs.packed, rs.readlens, k, m, 1, 65535, 0, via_supermers=False)
    got = po.canonicalize(k, *np.unique(w, axis=0, return_counts=True))
    po.assert_equal(got, po.Counts(k, exp.nwords, exp.words, exp.cnt, None, None, None), "supermer re-expansion")
    # ownership: every canonical k-mer lives in exactly one bucket
    uw, inv = np.unique(w, axis=0, return_inverse=True)
    first = np.full(len(uw), -1)
    first[inv] = b
    assert np.array_equal(first[inv], b)
    if ext:
        # (pos, rid) of every supermer point at its bases in the original read
        lens = sm["len"]
        for s in range(0, len(lens), 7):
            pos, rid = int(sm["ext"][s] >> np.uint64(32)), int(sm["ext"][s] & np.uint64(0xFFFFFFFF)) - 5
            codes = rs.codes(rid)[pos:pos + int(lens[s])]
            assert np.array_equal(codes, slot_codes(sm, s))


@pytest.mark.parametrize("k,n,with_val", [(31, 1, False), (31, 6143, False), (31, 6145, True), (31, 1_000_003, False),
                                          (55, 300_001, False), (55, 50_000, True), (90, 100_000, True), (15, 70_000, False)])
def test_radix_sort(k, n, with_val):
    import torch
    nw = 1 if k <= 32 else (2 if k <= 64 else 3)
    g = torch.Generator(device="cuda").manual_seed(k * 1000 + n)
    keys = []
    for w in range(nw):
        bases = 32 if w < nw - 1 else k - 32 * (nw - 1)
        x = torch.randint(-2**63, 2**63 - 1, (n,), dtype=torch.int64, device="cuda", generator=g)
        if bases < 32:
            x = x & ~((1 << (64 - 2 * bases)) - 1)
        if n > 1000:   # duplicates and skew
            x[: n // 3] = x[n // 3: 2 * (n // 3)]
            x[-(n // 10):] = x[0]
        keys.append(x)
    tmp = [torch.empty_like(x) for x in keys]
    val = torch.arange(n, dtype=torch.int64, device="cuda") if with_val else None
    vtmp = torch.empty_like(val) if with_val else None
    host = np.stack([x.cpu().numpy().view(np.uint64) for x in keys], axis=1)
    torch.cuda.synchronize()
    with capi.Context(k, min(k - 1, 17), 1, 50) as ctx:
        ctx.debug_sort([x.data_ptr() for x in keys], [x.data_ptr() for x in tmp], n, k,
                       val.data_ptr() if with_val else 0, vtmp.data_ptr() if with_val else 0)
    torch.cuda.synchronize()
    got = np.stack([x.cpu().numpy().view(np.uint64) for x in keys], axis=1)
    order = np.lexsort(tuple(host[:, w] for w in range(nw - 1, -1, -1)))
    assert np.array_equal(got, host[order])
    if with_val:
        v = val.cpu().numpy()
        assert np.array_equal(np.sort(v), np.arange(n))          # a permutation
        assert np.array_equal(host[v], got)                      # payload follows its key
        # LSD passes are stable: equal keys keep their input order
        same = np.all(got[1:] == got[:-1], axis=1)
        assert np.all(v[1:][same] > v[:-1][same])


def test_golden_fixtures(golden):
    g = golden
    with capi.Context(g["k"], g["m"], g["lower"], g["upper"], g["ext"]) as ctx:
        c, raw = gpu_counts(ctx, g["packed"], g["readlens"])
        po.assert_equal(c, g["expected"], "CUDA vs reference golden")
        assert c.histogram_text() == g["histogram_text"]
        hist = np.bincount(g["expected"].cnt.astype(np.int64), minlength=g["upper"] + 1).astype(np.uint64)
        assert np.array_equal(raw["histogram"], hist)
        assert np.array_equal(ctx.allreduce_histogram(), hist)
        if not g["ext"]:
            ent = ctx.fill_entries(raw["n_kept"])
            assert np.array_equal(ent[:, :-1], raw["words"]) and np.array_equal(ent[:, -1], raw["cnt"])


@pytest.mark.parametrize("k,m,ext,lower,upper", [(31, 17, 0, 2, 50), (55, 23, 0, 2, 50), (31, 17, 1, 2, 50),
                                                   (31, 17, 0, 1, 65535), (77, 40, 1, 1, 8), (17, 9, 0, 3, 12)])
@pytest.mark.parametrize("read_len", [150, 5000])
def test_against_oracle(k, m, ext, lower, upper, read_len):
    rs = synth.sample_fixed(120_000, 6.0, read_len, 0.01, seed=k + read_len)
    exp = po.kmer_count(rs.packed, rs.readlens, k, m, lower, upper, ext, via_supermers=False, readid_base=3)
    with capi.Context(k, m, lower, upper, ext) as ctx:
        c, raw = gpu_counts(ctx, rs.packed, rs.readlens, readid_base=3)
        po.assert_equal(c, exp, "CUDA vs oracle")
        assert raw["stats"]["n_kmers_local"] == rs.num_kmers(k)
        # idempotent and independent of batching / bucket count
        c2, _ = gpu_counts(ctx, rs.packed, rs.readlens, readid_base=3)
        po.assert_equal(c2, exp, "second call on the same context")
    # three huge bins: nothing fits on chip, everything takes the HBM path (expand -> radix sort -> count) in batches
    with capi.Context(k, m, lower, upper, ext, buckets_per_rank=3, batch_kmers=100_000) as ctx:
        c3, raw3 = gpu_counts(ctx, rs.packed, rs.readlens, readid_base=3)
        po.assert_equal(c3, exp, "HBM path, small batches")
        assert raw3["stats"]["n_overflow_bins"] == 3 and raw3["stats"]["n_batches"] >= 4
    # mid-sized bins: several supermer chunks per bin on chip (K <= 32 without EXT), overflow otherwise
    with capi.Context(k, m, lower, upper, ext, buckets_per_rank=37) as ctx:
        c4, raw4 = gpu_counts(ctx, rs.packed, rs.readlens, readid_base=3)
        po.assert_equal(c4, exp, "37 bins")


def test_bucket_balance():
    """uniform synthetic reads must spread evenly over the minimizer buckets (this is what balances
    the GPUs of a multi-rank run)."""
    rs = synth.sample_fixed(400_000, 10.0, 1000, 0.01, seed=21)
    for k, m in [(31, 17), (55, 23)]:
        with capi.Context(k, m, 1, 50, buckets_per_rank=64) as ctx:
            sm = ctx.debug_extract(rs.packed, rs.readlens)
        bk = sm["bucket_kmers"].astype(np.float64)
        assert bk.max() / bk.mean() < 1.5, (k, m, bk.max() / bk.mean())
        assert bk.min() / bk.mean() > 0.5


def test_edge_cases():
    k, m = 31, 17
    with capi.Context(k, m, 1, 50, 1) as ctx:
        # empty input
        rs = synth.pack_reads([])
        r = ctx.count(rs.packed, rs.readlens)
        assert r["n_kept"] == 0 and r["n_occ"] == 0
        # only reads shorter than K
        rs = synth.pack_reads([synth.ascii_to_codes("ACGT" * 7)] * 5)
        assert ctx.count(rs.packed, rs.readlens)["n_kept"] == 0
        # len == K, len % 4 in {0,1,2,3}, empty read in the middle, N -> A
        reads = [synth.ascii_to_codes("ACGT" * 7 + "ACG"), np.zeros(0, np.uint8), synth.ascii_to_codes("ACGT" * 8),
                 synth.ascii_to_codes("ACGTN" * 7), synth.ascii_to_codes("TTGCA" * 7 + "T"), synth.ascii_to_codes("G" * 34)]
        rs = synth.pack_reads(reads)
        exp = po.kmer_count(rs.packed, rs.readlens, k, m, 1, 50, 1, via_supermers=False)
        c, _ = gpu_counts(ctx, rs.packed, rs.readlens)
        po.assert_equal(c, exp, "edge reads")
    # counts exactly L, U, U+1; homopolymer runs longer than any supermer cap; count > 65535
    base = synth.make_genome(400, 5)
    reads = [base[:120]] * 3 + [base[150:300]] * 4 + [base[300:400]] * 5 + [synth.ascii_to_codes("A" * 5000)] * 20
    rs = synth.pack_reads(reads)
    for lower, upper in [(3, 4), (4, 4), (5, 5), (1, 65535), (2, 3)]:
        exp = po.kmer_count(rs.packed, rs.readlens, k, m, lower, upper, 0, via_supermers=False)
        with capi.Context(k, m, lower, upper, 0) as ctx:
            c, _ = gpu_counts(ctx, rs.packed, rs.readlens)
            po.assert_equal(c, exp, f"filter edges L={lower} U={upper}")


def test_errors():
    with pytest.raises(capi.HskError):
        capi.Context(2, 1, 1, 50)
    with pytest.raises(capi.HskError):
        capi.Context(31, 31, 1, 50)
    with pytest.raises(capi.HskError):
        capi.Context(31, 17, 0, 50)
    with pytest.raises(capi.HskError):
        capi.Context(31, 17, 2, 70000)
    with capi.Context(31, 17, 2, 50) as ctx:
        rs = synth.sample_fixed(1000, 2.0, 100, 0.0, seed=1)
        with pytest.raises(capi.HskError):   # buffer size inconsistent with the read lengths
            ctx.count(rs.packed[:-3], rs.readlens)


def test_large_properties():
    """BASELINE config[1] scale is covered by bench.py; here a 20 Mbp run checks size-independent
    properties: total of counts == N when nothing is filtered, histogram consistency, sortedness
    inside a batch, and equality with the oracle's direct definition."""
    k, m = 31, 17
    rs = synth.sample_fixed(1_000_000, 20.0, 2000, 0.01, seed=99)
    N = rs.num_kmers(k)
    with capi.Context(k, m, 1, 65535) as ctx:
        r = ctx.count(rs.packed, rs.readlens)
        assert int(r["cnt"].astype(np.uint64).sum()) == N
        assert int((r["histogram"] * np.arange(65536, dtype=np.uint64)).sum()) == N
        # the on-chip path takes (almost) every bin; entries ascend by k-mer inside a bin, so descents
        # only happen at bin boundaries
        nbins_est = 4 * rs.packed.nbytes // 2048
        assert r["stats"]["n_overflow_bins"] <= 0.01 * nbins_est + 2
        w = r["words"][:, 0]
        assert int((w[1:] < w[:-1]).sum()) <= nbins_est + 64
        exp = po.kmer_count(rs.packed, rs.readlens, k, m, 1, 65535, 0, via_supermers=False)
        po.assert_equal(po.canonicalize(k, r["words"], r["cnt"]), exp, "20 Mbp vs oracle")


def test_host_pipeline_groups_and_device_path_agree(monkeypatch):
    """hsk_count streams the arena out in groups of bins while the kernel runs (page-locked completion records);
    whatever the number of groups, the host result equals hsk_count_device + hsk_fetch_result and the oracle."""
    import torch
    k, m = 31, 17
    rs = synth.sample_fixed(600_000, 12.0, 3000, 0.01, seed=5)
    exp = po.kmer_count(rs.packed, rs.readlens, k, m, 2, 50, 0, via_supermers=False)
    off = rs.byte_offsets()
    d_packed = torch.zeros(((rs.packed.nbytes + 15) // 16) * 16 + 64, dtype=torch.uint8, device="cuda")
    d_packed[: rs.packed.nbytes].copy_(torch.from_numpy(rs.packed))
    d_off = torch.from_numpy(off.view(np.int64)).cuda()
    d_len = torch.from_numpy(rs.readlens.astype(np.uint32).view(np.int32)).cuda()
    torch.cuda.synchronize()
    with capi.Context(k, m, 2, 50, 0) as ctx:
        ctx.count_device(d_packed.data_ptr(), rs.packed.nbytes, d_off.data_ptr(), d_len.data_ptr(), rs.nreads)
        dev = ctx.fetch()
        po.assert_equal(po.canonicalize(k, dev["words"], dev["cnt"]), exp, "device path")
        for groups in ("1", "3", "8", "37"):
            monkeypatch.setenv("HSK_GROUPS", groups)
            r = ctx.count(rs.packed, rs.readlens)
            # same arena, same order: bins in index order, ascending k-mers inside a bin
            assert np.array_equal(r["words"], dev["words"]) and np.array_equal(r["cnt"], dev["cnt"]), groups
            assert np.array_equal(r["histogram"], dev["histogram"])
    monkeypatch.delenv("HSK_GROUPS")
    # EXTENSION through the same pipeline
    exp = po.kmer_count(rs.packed, rs.readlens, k, m, 2, 50, 1, via_supermers=False)
    with capi.Context(k, m, 2, 50, 1) as ctx:
        for groups in ("2", "8"):
            monkeypatch.setenv("HSK_GROUPS", groups)
            c, _ = gpu_counts(ctx, rs.packed, rs.readlens)
            po.assert_equal(c, exp, f"EXT, {groups} groups")


def test_read_table_many_short_and_empty_reads():
    """reads.cu: offsets / lengths of 50 000 reads (several scan tiles), every fourth one empty, lengths 0..99."""
    k, m = 21, 11
    rng = np.random.default_rng(3)
    genome = synth.make_genome(20_000, 9)
    reads = []
    for i in range(50_000):
        n = 0 if i % 4 == 0 else int(rng.integers(0, 100))
        st = int(rng.integers(0, len(genome) - 100))
        reads.append(genome[st:st + n])
    rs = synth.pack_reads(reads)
    exp = po.kmer_count(rs.packed, rs.readlens, k, m, 1, 65535, 1, via_supermers=False)
    with capi.Context(k, m, 1, 65535, 1) as ctx:
        c, raw = gpu_counts(ctx, rs.packed, rs.readlens)
        po.assert_equal(c, exp, "short and empty reads")
        assert raw["stats"]["n_kmers_local"] == rs.num_kmers(k)


@pytest.mark.parametrize("k,m,ext", [(31, 17, 0), (31, 17, 1), (55, 23, 0)])
def test_count_stream_parts_and_input_memory(k, m, ext, monkeypatch):
    """hsk_count_stream hands the result to a sink part by part while the kernel runs: the parts are contiguous, in
    order, their concatenation equals hsk_count's result and the oracle's; page-locked and pageable input (staged
    through the ring of page-locked buffers by the context's host threads) give the same result."""
    import torch
    rs = synth.sample_fixed(400_000, 25.0, 3000, 0.01, seed=31 + k + ext)
    exp = po.kmer_count(rs.packed, rs.readlens, k, m, 2, 50, ext, via_supermers=False)
    monkeypatch.setenv("HSK_GROUPS", "9")
    with capi.Context(k, m, 2, 50, ext) as ctx:
        base = ctx.count(rs.packed, rs.readlens)                      # pageable numpy memory
        s = ctx.count_stream(rs.packed, rs.readlens)
        parts = sorted(s["parts"])                                   # delivered by several threads, in any order
        firsts, ns = [p[0] for p in parts], [p[1] for p in parts]
        assert firsts == [int(x) for x in np.cumsum([0] + ns[:-1])] and sum(ns) == base["n_kept"] and len(ns) >= 2
        assert max(p[2] for p in parts) >= base["n_kept"] and min(p[2] for p in parts) >= 0.9 * base["n_kept"]   # the hints
        assert np.array_equal(s["words"], base["words"]) and np.array_equal(s["cnt"], base["cnt"])
        got = po.canonicalize(k, s["words"], s["cnt"], s.get("occ_off"), s.get("pos"), s.get("rid"))
        po.assert_equal(got, exp, "stream vs oracle")               # (the order inside an occurrence list is not fixed)
        if ext:
            assert np.array_equal(s["occ_off"], base["occ_off"])
        # page-locked input goes up from where it is
        hp = torch.from_numpy(rs.packed).pin_memory()
        hl = torch.from_numpy(rs.readlens.view(np.int64)).pin_memory()
        r = capi.Result()
        capi._check(ctx.lib.hsk_count(ctx.handle, hp.data_ptr(), rs.packed.nbytes, hl.data_ptr(), rs.nreads, 0, capi.C.byref(r)))
        pinned = ctx._unpack(r)
        assert np.array_equal(pinned["words"], base["words"]) and np.array_equal(pinned["cnt"], base["cnt"])
    # many small chunks through the staging ring: a 40 MB buffer is cut into 2 MB pieces
    rs2 = synth.sample_fixed(4_000_000, 10.0, 10_000, 0.005, seed=77)
    with capi.Context(31, 17, 2, 50, 0) as ctx:
        a = ctx.count(rs2.packed, rs2.readlens)
        hp = torch.from_numpy(rs2.packed).pin_memory()
        hl = torch.from_numpy(rs2.readlens.view(np.int64)).pin_memory()
        r = capi.Result()
        capi._check(ctx.lib.hsk_count(ctx.handle, hp.data_ptr(), rs2.packed.nbytes, hl.data_ptr(), rs2.nreads, 0, capi.C.byref(r)))
        b = ctx._unpack(r)
        assert a["n_kept"] == b["n_kept"] and np.array_equal(a["words"], b["words"]) and np.array_equal(a["cnt"], b["cnt"])
        assert int(a["stats"]["n_kmers_local"]) == rs2.num_kmers(31)


def test_arena_limits(monkeypatch):
    """Memory planning (engine.cu: count_device): when the arena is cut to what fits, a result that fits is unchanged, bins
    that find no room in the staging area take the HBM path, and a result that does not fit is an error, not a
    corrupted arena."""
    k, m = 31, 17
    rs = synth.sample_fixed(300_000, 12.0, 2000, 0.002, seed=8)
    exp = po.kmer_count(rs.packed, rs.readlens, k, m, 1, 65535, 0, via_supermers=False)
    # LOWER = 1 and few bins: every bin keeps more k-mers than a CTA sorts itself (staging area + big gather)
    with capi.Context(k, m, 1, 65535, 0, buckets_per_rank=128) as ctx:
        c, raw = gpu_counts(ctx, rs.packed, rs.readlens)
        po.assert_equal(c, exp, "big bins, roomy arena")
        assert raw["stats"]["n_overflow_bins"] == 0
    # 8 MB: arena of ~590 K entries (the result has ~320 K), staging area of ~74 K entries: most bins find it full
    monkeypatch.setenv("HSK_ARENA_BUDGET_MB", "8")
    with capi.Context(k, m, 1, 65535, 0, buckets_per_rank=128) as ctx:
        c, raw = gpu_counts(ctx, rs.packed, rs.readlens)
        po.assert_equal(c, exp, "big bins, staging area exhausted -> HBM path")
        assert raw["stats"]["n_overflow_bins"] > 0
        c2, _ = gpu_counts(ctx, rs.packed, rs.readlens)
        po.assert_equal(c2, exp, "second call")
    # EXTENSION with a tight budget
    exp1 = po.kmer_count(rs.packed, rs.readlens, k, m, 2, 50, 1, via_supermers=False)
    monkeypatch.setenv("HSK_ARENA_BUDGET_MB", "64")
    with capi.Context(k, m, 2, 50, 1) as ctx:
        c, _ = gpu_counts(ctx, rs.packed, rs.readlens)
        po.assert_equal(c, exp1, "EXT, capped arena")
    # 1 MB: the arena cannot hold the result
    monkeypatch.setenv("HSK_ARENA_BUDGET_MB", "1")
    with capi.Context(k, m, 1, 65535, 0) as ctx:
        with pytest.raises(capi.HskError, match="does not fit"):
            ctx.count(rs.packed, rs.readlens)
        monkeypatch.delenv("HSK_ARENA_BUDGET_MB")
        c, _ = gpu_counts(ctx, rs.packed, rs.readlens)               # the context is usable afterwards
        po.assert_equal(c, exp, "after the error")


def test_heavy_bin_poly_a():
    """One minimizer bin far above the others (the reference pre-counts such tasks, kmerops.cpp:1157-1199): 60 Mbp of
    poly-A among ordinary reads.  The bin is counted on chip by one CTA (one distinct k-mer), its total exceeds 2^24
    supermer slots only at larger sizes — here the per-bin totals are checked against pass A's independent totals."""
    k, m = 31, 17
    rs = synth.sample_fixed(200_000, 8.0, 2000, 0.01, seed=4)
    polya = np.zeros((6000, 2500), dtype=np.uint8)   # 6000 reads x 10 000 A
    reads_packed = np.concatenate([rs.packed, polya.reshape(-1)])
    lens = np.concatenate([rs.readlens, np.full(6000, 10_000, dtype=np.uint64)])
    with capi.Context(k, m, 2, 65535, 0) as ctx:
        r = ctx.count(reads_packed, lens)
        assert r["stats"]["n_kmers_local"] == rs.num_kmers(k) + 6000 * (10_000 - k + 1)
        exp = po.kmer_count(rs.packed, rs.readlens, k, m, 2, 65535, 0, via_supermers=False)
        got = po.canonicalize(k, r["words"], r["cnt"])
        # AAAA...A occurs ~6e7 times: above UPPER, dropped; everything else as without the poly-A reads (a genomic
        # poly-A k-mer would have been counted with them, a uniform 200 kbp genome has none)
        po.assert_equal(got, exp, "poly-A poisoned input")
