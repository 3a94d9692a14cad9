"""ctypes binding of the hysortk C++ API (include/hysortk.hpp) through the C entry points of
hysortk_b200/cxx/bench_api.cpp: what a HySortK caller links — ``hysortk::kmer_count(const DnaBuffer&, MPI_Comm)``
returning a ``std::vector<KmerListEntryS>`` — called from Python by the tests and by bench.py's end-to-end leg.

One shared library per compile-time configuration (the reference's ``make K= M= L= U= EXT=`` parameters), built in
``hysortk_b200/_api/`` on top of ``libhysortk_b200.so`` (the CUDA engine): only the C++ host layer is compiled per
configuration.  Several ranks: set HSK_MPI_SIZE / HSK_MPI_RANK / HSK_MPI_SESSION (bundled MPI stand-in) before the
first call; LOCAL_RANK picks the GPU.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import build as _build

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
APIDIR = os.path.join(HERE, "_api")
CXX_SOURCES = ["hysortk.cpp", "dnaseq.cpp", "dnabuffer.cpp", "hashfuncs.cpp", "bench_api.cpp"]


def so_path(k: int, m: int, lower: int, upper: int, ext: int) -> str:
    return os.path.join(APIDIR, f"libhysortk_api_k{k}_m{m}_l{lower}_u{upper}_e{ext}.so")


def build(k: int, m: int, lower: int, upper: int, ext: int, force: bool = False) -> str:
    """g++ of the C++ host layer for one configuration, linked against the engine library next to it."""
    engine = _build.build()
    so = so_path(k, m, lower, upper, ext)
    srcs = [os.path.join(HERE, "cxx", f) for f in CXX_SOURCES]
    deps = srcs + [engine] + [os.path.join(ROOT, "include", f) for f in os.listdir(os.path.join(ROOT, "include"))] + \
        [os.path.join(HERE, "shim", "mpi.h")]
    if not force and os.path.exists(so) and os.path.getmtime(so) >= max(os.path.getmtime(p) for p in deps):
        return so
    os.makedirs(APIDIR, exist_ok=True)
    mpi_inc = os.environ.get("HSK_MPI_INC", "-I" + os.path.join(HERE, "shim"))
    cmd = ["g++", "-O3", "-std=c++17", "-fPIC", "-shared", "-fopenmp", "-pthread", "-Wall",
           f"-DKMER_SIZE={k}", f"-DMINIMIZER_SIZE={m}", f"-DLOWER_KMER_FREQ={lower}", f"-DUPPER_KMER_FREQ={upper}",
           f"-DEXTENSION={ext}", "-DLOG_LEVEL=0", "-DDEBUG=0", "-I" + os.path.join(ROOT, "include"), mpi_inc, *srcs,
           "-L" + HERE, "-lhysortk_b200", "-Wl,-rpath,$ORIGIN/..", "-lrt", "-o", so]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"building {so} failed:\n{r.stdout}\n{r.stderr}")
    return so


class ApiError(RuntimeError):
    pass


_LIBS: dict = {}


def load(k: int, m: int, lower: int, upper: int, ext: int):
    key = (k, m, lower, upper, ext)
    if key not in _LIBS:
        so = so_path(*key)
        if not os.path.exists(so):
            so = build(*key)
        L = C.CDLL(so)
        L.hsk_api_last_error.restype = C.c_char_p
        L.hsk_api_kmer_count.restype = C.c_void_p
        L.hsk_api_kmer_count.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.hsk_api_kmer_count_fasta.restype = C.c_void_p
        L.hsk_api_kmer_count_fasta.argtypes = [C.c_char_p]
        L.hsk_api_bench.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p]
        L.hsk_api_seconds.restype = C.c_double
        L.hsk_api_seconds.argtypes = [C.c_void_p]
        L.hsk_api_size.restype = C.c_size_t
        L.hsk_api_size.argtypes = [C.c_void_p]
        L.hsk_api_total_occurrences.restype = C.c_size_t
        L.hsk_api_total_occurrences.argtypes = [C.c_void_p]
        L.hsk_api_export.argtypes = [C.c_void_p] * 6
        L.hsk_api_print_histogram.argtypes = [C.c_void_p, C.c_char_p]
        L.hsk_api_write_output.argtypes = [C.c_void_p, C.c_char_p]
        L.hsk_api_free.argtypes = [C.c_void_p]
        L.hsk_api_set_threads.argtypes = [C.c_int]
        pk = [C.c_int() for _ in range(6)]
        L.hsk_api_params(*[C.byref(x) for x in pk])
        assert tuple(x.value for x in pk[:5]) == key, "library built for another configuration"
        _LIBS[key] = L
    return _LIBS[key]


def kmer_count(packed: np.ndarray, readlens: np.ndarray, k: int, m: int, lower: int, upper: int, ext: int = 0,
               fasta: str | None = None, want_text_dir: str | None = None) -> dict:
    """hysortk::kmer_count through the C++ API; returns the entries of the returned KmerListS as arrays (in list order)
    and the wall time of the call.  With want_text_dir also print_kmer_histogram / write_output_file into that directory
    (hist.txt, <rank>.out)."""
    L = load(k, m, lower, upper, ext)
    if fasta is not None:
        h = L.hsk_api_kmer_count_fasta(fasta.encode())
    else:
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        rl = np.ascontiguousarray(readlens, dtype=np.uint64)   # size_t
        h = L.hsk_api_kmer_count(packed.ctypes.data, packed.nbytes, rl.ctypes.data, len(rl))
    if not h:
        raise ApiError(L.hsk_api_last_error().decode(errors="replace"))
    try:
        n = L.hsk_api_size(h)
        nw = 1 if k <= 32 else (2 if k <= 64 else 3)
        words = np.zeros(max(n, 1) * nw, dtype=np.uint64)
        cnt = np.zeros(max(n, 1), dtype=np.uint64)
        out = dict(n=n, seconds=L.hsk_api_seconds(h))
        if ext:
            tot = L.hsk_api_total_occurrences(h)
            occ_off = np.zeros(n + 1, dtype=np.uint64)
            pos = np.zeros(max(tot, 1), dtype=np.uint32)
            rid = np.zeros(max(tot, 1), dtype=np.int32)
            L.hsk_api_export(h, words.ctypes.data, cnt.ctypes.data, occ_off.ctypes.data, pos.ctypes.data, rid.ctypes.data)
            out.update(occ_off=occ_off, pos=pos[:tot], rid=rid[:tot])
        else:
            L.hsk_api_export(h, words.ctypes.data, cnt.ctypes.data, None, None, None)
        out.update(words=words[: n * nw].reshape(n, nw), cnt=cnt[:n])
        if want_text_dir:
            L.hsk_api_print_histogram(h, os.path.join(want_text_dir, "hist.txt").encode())
            L.hsk_api_write_output(h, want_text_dir.encode())
        return out
    finally:
        L.hsk_api_free(h)


def bench(packed: np.ndarray, readlens: np.ndarray, k: int, m: int, lower: int, upper: int, ext: int, warmup: int, steps: int,
          threads: int | None = None) -> dict:
    """`steps` timed calls of hysortk::kmer_count on one pageable DnaBuffer (this rank's share)."""
    L = load(k, m, lower, upper, ext)
    if threads:
        L.hsk_api_set_threads(int(threads))
    packed = np.ascontiguousarray(packed, dtype=np.uint8)
    rl = np.ascontiguousarray(readlens, dtype=np.uint64)
    secs = np.zeros(max(steps, 1), dtype=np.float64)
    kept, occ, chk = C.c_uint64(), C.c_uint64(), C.c_uint64()
    rc = L.hsk_api_bench(packed.ctypes.data, packed.nbytes, rl.ctypes.data, len(rl), warmup, steps, secs.ctypes.data,
                         C.byref(kept), C.byref(occ), C.byref(chk))
    if rc:
        raise ApiError(L.hsk_api_last_error().decode(errors="replace"))
    return dict(seconds=secs[:steps].copy(), n_kept=int(kept.value), n_occ=int(occ.value), checksum=int(chk.value))


def release(k: int, m: int, lower: int, upper: int, ext: int) -> None:
    load(k, m, lower, upper, ext).hsk_api_release()
