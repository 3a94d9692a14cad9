"""TEST INFRASTRUCTURE — ctypes bindings to the parity oracle.

* ``oracle.c``: our plain-C restatement of the reference's kmer_count path (liboracle.so).
* ``oracle/_ref/libhysortk_ref_*.so``: the UNMODIFIED reference compiled by ``build_ref.sh``
  (present in this container and on the GPU box as a prebuilt file; absent from git history).

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build_oracle(force: bool = False) -> str:
    so = os.path.join(HERE, "liboracle.so")
    src = os.path.join(HERE, "oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src),
                                                                    os.path.getmtime(os.path.join(HERE, "oracle.h"))):
        subprocess.check_call(["gcc", "-O2", "-std=c11", "-shared", "-fPIC", "-Wall", "-o", so, src])
    return so


class _OrcResult(C.Structure):
    _fields_ = [("k", C.c_int), ("m", C.c_int), ("lower", C.c_int), ("upper", C.c_int), ("ext", C.c_int),
                ("nwords", C.c_int), ("total_kmers", C.c_uint64), ("n_supermers", C.c_uint64),
                ("supermer_bytes", C.c_uint64), ("n", C.c_uint64), ("words", C.POINTER(C.c_uint64)),
                ("cnt", C.POINTER(C.c_uint64)), ("occ_off", C.POINTER(C.c_uint64)), ("pos", C.POINTER(C.c_uint32)),
                ("rid", C.POINTER(C.c_int32)), ("hist_len", C.c_uint64), ("hist", C.POINTER(C.c_uint64))]


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build_oracle())
        L.orc_kmer_count.restype = C.POINTER(_OrcResult)
        L.orc_kmer_count.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_int, C.c_int32, C.c_int]
        L.orc_free.argtypes = [C.POINTER(_OrcResult)]
        L.orc_histogram_text.restype = C.c_size_t
        L.orc_histogram_text.argtypes = [C.POINTER(_OrcResult), C.c_char_p, C.c_size_t]
        L.orc_output_text.restype = C.c_size_t
        L.orc_output_text.argtypes = [C.POINTER(_OrcResult), C.c_char_p, C.c_size_t]
        L.orc_murmur3_64.restype = C.c_uint64
        L.orc_murmur3_64.argtypes = [C.c_void_p, C.c_uint32]
        L.orc_read_destinations.restype = C.c_size_t
        L.orc_read_destinations.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.orc_pack_read.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p]
        L.orc_kmer_set.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
        L.orc_kmer_twin.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_kmer_rep.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_kmer_extend.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.orc_kmer_string.argtypes = [C.c_void_p, C.c_int, C.c_char_p]
        _LIB = L
    return _LIB


@dataclass
class Counts:
    """Canonical form of a k-mer count result: k-mers ascending by Kmer::operator< (word 0 most
    significant); for EXT the occurrences of each k-mer ascending by (rid, pos)."""

    k: int
    nwords: int
    words: np.ndarray           # (n, nwords) uint64
    cnt: np.ndarray             # (n,) uint64
    occ_off: np.ndarray | None  # (n+1,) uint64
    pos: np.ndarray | None      # uint32
    rid: np.ndarray | None      # int32
    hist: np.ndarray | None = None  # (upper+1,) uint64
    total_kmers: int = 0
    seconds: float = 0.0
    extra: dict | None = None

    @property
    def n(self) -> int:
        return int(self.cnt.shape[0])

    def histogram(self, upper: int) -> np.ndarray:
        return np.bincount(self.cnt.astype(np.int64), minlength=upper + 1).astype(np.uint64)

    def histogram_text(self) -> str:
        """reference src/hysortk.cpp:98-136"""
        h = np.bincount(self.cnt.astype(np.int64)) if self.n else np.zeros(1, dtype=np.int64)
        lines = ["#count\tnumkmers"]
        lines += [f"{i}\t{int(h[i])}" for i in range(1, len(h)) if h[i] > 0]
        return "\n".join(lines) + "\n\n"

    def strings(self) -> list[str]:
        out = []
        for row in self.words:
            s = []
            for i in range(self.k):
                s.append("ACGT"[(int(row[i // 32]) >> (2 * (31 - i % 32))) & 3])
            out.append("".join(s))
        return out


def canonicalize(k: int, words: np.ndarray, cnt: np.ndarray, occ_off=None, pos=None, rid=None, **kw) -> Counts:
    """Sort an arbitrary-order result into the canonical comparison order."""
    nw = 1 if k <= 32 else (2 if k <= 64 else 3)
    words = np.asarray(words, dtype=np.uint64).reshape(-1, nw)
    cnt = np.asarray(cnt).astype(np.uint64)
    order = np.lexsort(tuple(words[:, w] for w in range(nw - 1, -1, -1))) if len(cnt) else np.zeros(0, dtype=np.int64)
    w2, c2 = words[order], cnt[order]
    if occ_off is None:
        return Counts(k, nw, w2, c2, None, None, None, **kw)
    occ_off = np.asarray(occ_off, dtype=np.uint64)
    pos = np.asarray(pos, dtype=np.uint32)
    rid = np.asarray(rid, dtype=np.int32)
    lens = (occ_off[1:] - occ_off[:-1]).astype(np.int64)
    new_lens = lens[order]
    new_off = np.zeros(len(order) + 1, dtype=np.uint64)
    np.cumsum(new_lens, out=new_off[1:])
    # gather occurrences entry by entry, then sort inside each entry by (rid, pos)
    total = int(new_off[-1])
    src_start = occ_off[:-1][order].astype(np.int64)
    idx = np.repeat(src_start - new_off[:-1].astype(np.int64), new_lens) + np.arange(total, dtype=np.int64)
    p2, r2 = pos[idx], rid[idx]
    entry = np.repeat(np.arange(len(order), dtype=np.int64), new_lens)
    o2 = np.lexsort((p2, r2, entry))
    return Counts(k, nw, w2, c2, new_off, p2[o2], r2[o2], **kw)


def assert_equal(a: Counts, b: Counts, what: str = "") -> None:
    assert a.n == b.n, f"{what}: #kept k-mers differ: {a.n} vs {b.n}"
    assert np.array_equal(a.words, b.words), f"{what}: k-mer words differ"
    assert np.array_equal(a.cnt, b.cnt), f"{what}: counts differ"
    if a.occ_off is not None or b.occ_off is not None:
        assert a.occ_off is not None and b.occ_off is not None, f"{what}: extension info missing on one side"
        assert np.array_equal(a.occ_off, b.occ_off), f"{what}: occurrence offsets differ"
        assert np.array_equal(a.rid, b.rid), f"{what}: ReadIds differ"
        assert np.array_equal(a.pos, b.pos), f"{what}: PosInRead differ"


def kmer_count(packed: np.ndarray, readlens: np.ndarray, k: int, m: int, lower: int, upper: int, ext: int = 0,
               ntasks: int = 5, readid_base: int = 0, via_supermers: bool = True) -> Counts:
    L = lib()
    packed = np.ascontiguousarray(packed, dtype=np.uint8)
    readlens = np.ascontiguousarray(readlens, dtype=np.uint64)
    r = L.orc_kmer_count(packed.ctypes.data, readlens.ctypes.data, len(readlens), k, m, lower, upper, ext, ntasks,
                         readid_base, 1 if via_supermers else 0)
    try:
        rr = r.contents
        n, nw = int(rr.n), int(rr.nwords)
        words = np.ctypeslib.as_array(rr.words, shape=(max(n, 1) * nw,))[: n * nw].copy().reshape(n, nw)
        cnt = np.ctypeslib.as_array(rr.cnt, shape=(max(n, 1),))[:n].copy()
        hist = np.ctypeslib.as_array(rr.hist, shape=(int(rr.hist_len),)).copy()
        occ_off = pos = rid = None
        if ext:
            occ_off = np.ctypeslib.as_array(rr.occ_off, shape=(n + 1,)).copy()
            tot = int(occ_off[-1])
            pos = np.ctypeslib.as_array(rr.pos, shape=(max(tot, 1),))[:tot].copy()
            rid = np.ctypeslib.as_array(rr.rid, shape=(max(tot, 1),))[:tot].copy()
        return Counts(k, nw, words, cnt, occ_off, pos, rid, hist=hist, total_kmers=int(rr.total_kmers),
                      extra=dict(n_supermers=int(rr.n_supermers), supermer_bytes=int(rr.supermer_bytes)))
    finally:
        L.orc_free(r)


def output_text(c: Counts) -> str:
    """reference src/hysortk.cpp:149-162"""
    return "".join(f"{s}\t{int(n)}\n" for s, n in zip(c.strings(), c.cnt))


# ----------------------------------------------------------------------------- the real reference

def ref_so(k: int, m: int, lower: int, upper: int, ext: int) -> str:
    return os.path.join(HERE, "_ref", f"libhysortk_ref_k{k}_m{m}_l{lower}_u{upper}_e{ext}.so")


def ref_available(k: int, m: int, lower: int, upper: int, ext: int) -> bool:
    return os.path.exists(ref_so(k, m, lower, upper, ext))


_REF_LIBS: dict = {}


def ref_lib(k: int, m: int, lower: int, upper: int, ext: int):
    key = (k, m, lower, upper, ext)
    if key not in _REF_LIBS:
        L = C.CDLL(ref_so(*key))
        L.ref_kmer_count.restype = C.c_void_p
        L.ref_kmer_count.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.ref_kmer_count_fasta.restype = C.c_void_p
        L.ref_kmer_count_fasta.argtypes = [C.c_char_p]
        L.ref_seconds.restype = C.c_double
        L.ref_seconds.argtypes = [C.c_void_p]
        L.ref_size.restype = C.c_size_t
        L.ref_size.argtypes = [C.c_void_p]
        L.ref_total_occurrences.restype = C.c_size_t
        L.ref_total_occurrences.argtypes = [C.c_void_p]
        L.ref_export.argtypes = [C.c_void_p] * 6
        L.ref_print_histogram.argtypes = [C.c_void_p, C.c_char_p]
        L.ref_write_output.argtypes = [C.c_void_p, C.c_char_p]
        L.ref_free.argtypes = [C.c_void_p]
        pk = [C.c_int() for _ in range(6)]
        L.ref_params(*[C.byref(x) for x in pk])
        assert (pk[0].value, pk[1].value, pk[2].value, pk[3].value, pk[4].value) == key
        _REF_LIBS[key] = L
    return _REF_LIBS[key]


def ref_kmer_count(packed: np.ndarray, readlens: np.ndarray, k: int, m: int, lower: int, upper: int, ext: int = 0,
                   fasta: str | None = None, want_text: bool = False) -> Counts:
    """Runs the reference's own kmer_count (single rank, all OpenMP threads)."""
    L = ref_lib(k, m, lower, upper, ext)
    if fasta is not None:
        h = L.ref_kmer_count_fasta(fasta.encode())
    else:
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        rl = np.ascontiguousarray(readlens, dtype=np.uint64)  # size_t
        h = L.ref_kmer_count(packed.ctypes.data, packed.nbytes, rl.ctypes.data, len(rl))
    try:
        n = L.ref_size(h)
        nw = 1 if k <= 32 else (2 if k <= 64 else 3)
        words = np.zeros(max(n, 1) * nw, dtype=np.uint64)
        cnt = np.zeros(max(n, 1), dtype=np.uint64)
        occ_off = pos = rid = None
        if ext:
            tot = L.ref_total_occurrences(h)
            occ_off = np.zeros(n + 1, dtype=np.uint64)
            pos = np.zeros(max(tot, 1), dtype=np.uint32)
            rid = np.zeros(max(tot, 1), dtype=np.int32)
            L.ref_export(h, words.ctypes.data, cnt.ctypes.data, occ_off.ctypes.data, pos.ctypes.data, rid.ctypes.data)
            pos, rid = pos[:tot], rid[:tot]
        else:
            L.ref_export(h, words.ctypes.data, cnt.ctypes.data, None, None, None)
        extra = {}
        if want_text:
            with tempfile.TemporaryDirectory() as d:
                L.ref_print_histogram(h, os.path.join(d, "hist.txt").encode())
                L.ref_write_output(h, d.encode())
                extra["histogram_text"] = open(os.path.join(d, "hist.txt")).read()
                extra["output_text_raw"] = open(os.path.join(d, "0.out")).read()
        c = canonicalize(k, words[: n * nw], cnt[:n], occ_off, pos, rid, seconds=L.ref_seconds(h))
        c.extra = extra
        return c
    finally:
        L.ref_free(h)


# ------------------------------------------------------------- the real reference, several MPI ranks

def _partition(readlens: np.ndarray, nranks: int) -> np.ndarray:
    """first read of every rank: the reference's greedy contiguous partition by bases
    (reference src/fastaindex.cpp:52-100)."""
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
                sofar += int(lens[rid]); rid += 1
                if not (rid < n and sofar + int(lens[rid]) < avg):
                    break
    first[nranks - 1] = rid
    return first


def ref_kmer_count_ranks(packed: np.ndarray, readlens: np.ndarray, k: int, m: int, lower: int, upper: int, ext: int = 0,
                         nranks: int = 2, threads_per_rank: int | None = None, want_result: bool = True,
                         repeats: int = 1, timeout: float = 1800.0):
    """The reference's kmer_count as an MPI job of `nranks` processes of this node (the bundled multi-process MPI
    stand-in, hysortk_b200/shim/mpi.h): every rank gets its contiguous share of the reads, like read_dna_buffer
    would give it.  Returns (Counts of the union over ranks or None, list of per-repeat seconds = max over ranks)."""
    import json
    import sys
    import uuid
    packed = np.ascontiguousarray(packed, dtype=np.uint8)
    rl = np.ascontiguousarray(readlens, dtype=np.uint64)
    first = _partition(rl, nranks)
    nb = (rl + np.uint64(3)) // np.uint64(4)
    off = np.zeros(len(rl) + 1, dtype=np.uint64)
    np.cumsum(nb, out=off[1:])
    session = "ref" + uuid.uuid4().hex[:12]
    threads = threads_per_rank or max(1, (os.cpu_count() or 1) // nranks)
    with tempfile.TemporaryDirectory() as d:
        procs = []
        for r in range(nranks):
            lo, hi = int(first[r]), int(first[r + 1])
            np.savez(os.path.join(d, f"in{r}.npz"), packed=packed[int(off[lo]):int(off[hi])], readlens=rl[lo:hi])
            env = dict(os.environ, HSK_MPI_SIZE=str(nranks), HSK_MPI_RANK=str(r), HSK_MPI_SESSION=session,
                       OMP_NUM_THREADS=str(threads), OMP_PROC_BIND="false", SLURM_TASKS_PER_NODE=str(nranks))
            cmd = [sys.executable, os.path.join(HERE, "ref_worker.py"), os.path.join(d, f"in{r}.npz"),
                   os.path.join(d, f"out{r}.npz"), str(k), str(m), str(lower), str(upper), str(ext), str(repeats),
                   "1" if want_result else "0"]
            procs.append(subprocess.Popen(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
        outs = []
        try:
            for p in procs:
                o, e = p.communicate(timeout=timeout)
                if p.returncode != 0:
                    raise RuntimeError(f"reference rank failed (rc {p.returncode}):\n{o[-2000:]}\n{e[-2000:]}")
                outs.append(json.loads(o.strip().splitlines()[-1]))
        finally:
            for p in procs:
                if p.poll() is None:
                    p.kill()
            try:
                os.unlink("/dev/shm/hsk_mpi_" + session)
            except OSError:
                pass
        seconds = [max(o["seconds"][i] for o in outs) for i in range(repeats)]
        if not want_result:
            return None, seconds
        nw = 1 if k <= 32 else (2 if k <= 64 else 3)
        zs = [np.load(os.path.join(d, f"out{r}.npz")) for r in range(nranks)]
        words = np.concatenate([z["words"].reshape(-1, nw) for z in zs])
        cnt = np.concatenate([z["cnt"] for z in zs])
        if ext:
            offs, shift = [np.zeros(1, dtype=np.uint64)], 0
            for z in zs:
                offs.append(z["occ_off"][1:] + np.uint64(shift))
                shift += int(z["occ_off"][-1])
            c = canonicalize(k, words, cnt, np.concatenate(offs), np.concatenate([z["pos"] for z in zs]),
                             np.concatenate([z["rid"] for z in zs]))
        else:
            c = canonicalize(k, words, cnt)
        return c, seconds
