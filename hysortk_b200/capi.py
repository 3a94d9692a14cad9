"""ctypes binding of the C ABI in include/hsk_capi.h (libhysortk_b200.so).

This is the same boundary the C++ shim (include/hysortk.hpp, hysortk_b200/cxx/hysortk.cpp) calls;
tests and bench.py go through it so that what they exercise is what a HySortK caller links.
There is no CPU fallback: if the CUDA library is missing or no sm_100 device is present the calls
raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(HERE, "libhysortk_b200.so")
NCCL_ID_BYTES = 128

EXPORTS = ["hsk_last_error", "hsk_version", "hsk_get_unique_id", "hsk_create", "hsk_destroy", "hsk_count",
           "hsk_count_stream", "hsk_count_device", "hsk_fetch_result", "hsk_host_register", "hsk_host_unregister", "hsk_allreduce_histogram", "hsk_fill_entries", "hsk_debug_sort",
           "hsk_debug_extract"]


class HskError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [("k", C.c_int32), ("m", C.c_int32), ("lower", C.c_int32), ("upper", C.c_int32), ("ext", C.c_int32),
                ("device", C.c_int32), ("rank", C.c_int32), ("nranks", C.c_int32), ("nccl_id", C.c_void_p),
                ("buckets_per_rank", C.c_int32), ("batch_kmers", C.c_uint64), ("stream", C.c_void_p)]


class Stats(C.Structure):
    _fields_ = [("n_kmers_local", C.c_uint64), ("n_kmers_owned", C.c_uint64), ("n_supermers", C.c_uint64),
                ("supermer_bytes", C.c_uint64), ("bytes_sent", C.c_uint64), ("bytes_received", C.c_uint64),
                ("n_batches", C.c_uint64), ("n_sort_passes", C.c_uint64), ("n_launches", C.c_uint64),
                ("ms_h2d", C.c_float), ("ms_extract", C.c_float), ("ms_exchange", C.c_float), ("ms_expand", C.c_float),
                ("ms_sort", C.c_float), ("ms_count", C.c_float), ("ms_d2h", C.c_float), ("ms_total", C.c_float),
                ("ms_sort_passes", C.c_float), ("ms_bins", C.c_float), ("n_overflow_bins", C.c_uint64)]

    def as_dict(self) -> dict:
        return {n: getattr(self, n) for n, _ in self._fields_}


class Result(C.Structure):
    _fields_ = [("nwords", C.c_int32), ("n_kept", C.c_uint64), ("n_occ", C.c_uint64),
                ("kmer_words", C.POINTER(C.c_uint64)), ("cnt", C.POINTER(C.c_uint32)),
                ("occ_off", C.POINTER(C.c_uint64)), ("pos", C.POINTER(C.c_uint32)), ("rid", C.POINTER(C.c_int32)),
                ("histogram", C.POINTER(C.c_uint64)), ("stats", Stats)]


class DeviceResult(C.Structure):
    _fields_ = [("nwords", C.c_int32), ("n_kept", C.c_uint64), ("n_occ", C.c_uint64), ("d_kmer_words", C.c_void_p),
                ("d_cnt", C.c_void_p), ("d_occ_off", C.c_void_p), ("d_pos", C.c_void_p), ("d_rid", C.c_void_p),
                ("d_histogram", C.c_void_p), ("stats", Stats)]


# hsk_sink_fn: int sink(void *user, const hsk_result *view, first_entry, n_entries, first_occ, n_occ, total_hint)
SINK_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(Result), C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64)


class Supermers(C.Structure):
    _fields_ = [("n_bins", C.c_uint64), ("bin_slots", C.POINTER(C.c_uint64)), ("bin_kmers", C.POINTER(C.c_uint64)),
                ("n_slots", C.c_uint64), ("slot_words", C.c_uint32), ("slots", C.POINTER(C.c_uint32))]


_LIB = None


def load():
    """Loads libhysortk_b200.so (raises if it has not been built: no fallback)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(SO_PATH):
            raise HskError(f"{SO_PATH} not built; run `python -m hysortk_b200.build` (needs nvcc)")
        L = C.CDLL(SO_PATH)
        L.hsk_last_error.restype = C.c_char_p
        L.hsk_version.restype = C.c_int
        L.hsk_get_unique_id.argtypes = [C.c_void_p]
        L.hsk_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(Config)]
        L.hsk_destroy.argtypes = [C.c_void_p]
        L.hsk_destroy.restype = None
        L.hsk_count.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_int32, C.POINTER(Result)]
        L.hsk_count_stream.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_int32, C.c_void_p,
                                       C.c_void_p, C.POINTER(Result)]
        L.hsk_count_device.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int32,
                                       C.POINTER(DeviceResult)]
        L.hsk_fetch_result.argtypes = [C.c_void_p, C.POINTER(Result)]
        L.hsk_allreduce_histogram.argtypes = [C.c_void_p, C.c_void_p]
        L.hsk_fill_entries.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        L.hsk_debug_sort.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int32,
                                     C.c_int32]
        L.hsk_debug_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_int32,
                                        C.POINTER(Supermers)]
        _LIB = L
    return _LIB


def _check(rc: int) -> None:
    if rc != 0:
        raise HskError(load().hsk_last_error().decode(errors="replace"))


def get_unique_id() -> bytes:
    buf = C.create_string_buffer(NCCL_ID_BYTES)
    _check(load().hsk_get_unique_id(buf))
    return buf.raw


def _arr(ptr, n, dtype):
    if n == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(int(n),)).astype(dtype, copy=True)


class Context:
    """One engine context = one rank on one GPU (hsk_create / hsk_destroy)."""

    def __init__(self, k: int, m: int, lower: int, upper: int, ext: int = 0, device: int = 0, rank: int = 0,
                 nranks: int = 1, nccl_id: bytes | None = None, buckets_per_rank: int = 0, batch_kmers: int = 0,
                 stream: int | None = None):
        self.lib = load()
        self.k, self.m, self.lower, self.upper, self.ext = k, m, lower, upper, ext
        self.rank, self.nranks = rank, nranks
        self.nwords = 1 if k <= 32 else (2 if k <= 64 else 3)
        self._id = C.create_string_buffer(nccl_id, NCCL_ID_BYTES) if nccl_id is not None else None
        cfg = Config(k, m, lower, upper, ext, device, rank, nranks,
                     C.cast(self._id, C.c_void_p) if self._id is not None else None, buckets_per_rank, batch_kmers,
                     C.c_void_p(stream) if stream else None)
        self.handle = C.c_void_p()
        _check(self.lib.hsk_create(C.byref(self.handle), C.byref(cfg)))

    def close(self) -> None:
        if self.handle:
            self.lib.hsk_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- hot path ------------------------------------------------------------------------------------
    def _unpack(self, r: Result) -> dict:
        n, nw = int(r.n_kept), int(r.nwords)
        out = dict(nwords=nw, n_kept=n, n_occ=int(r.n_occ),
                   words=_arr(r.kmer_words, n * nw, np.uint64).reshape(n, nw), cnt=_arr(r.cnt, n, np.uint32),
                   histogram=_arr(r.histogram, self.upper + 1, np.uint64), stats=r.stats.as_dict())
        if self.ext:
            out["occ_off"] = _arr(r.occ_off, n + 1, np.uint64)
            out["pos"] = _arr(r.pos, int(r.n_occ), np.uint32)
            out["rid"] = _arr(r.rid, int(r.n_occ), np.int32)
        return out

    def count(self, packed: np.ndarray, readlens: np.ndarray, readid_base: int = 0, copy: bool = True) -> dict:
        """hsk_count on host buffers (H2D / D2H inside the call)."""
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        readlens = np.ascontiguousarray(readlens, dtype=np.uint64)
        r = Result()
        _check(self.lib.hsk_count(self.handle, packed.ctypes.data, packed.nbytes, readlens.ctypes.data, len(readlens),
                                  readid_base, C.byref(r)))
        if not copy:
            return dict(n_kept=int(r.n_kept), n_occ=int(r.n_occ), stats=r.stats.as_dict())
        return self._unpack(r)

    def count_stream(self, packed: np.ndarray, readlens: np.ndarray, readid_base: int = 0) -> dict:
        """hsk_count_stream: the result is collected part by part by a sink (a Python callback, called from the
        context's delivery threads, in any order) while the GPU is still counting; returns the parts put together like
        count() plus `parts` = the (first_entry, n_entries, total_hint) of every delivery, in arrival order."""
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        readlens = np.ascontiguousarray(readlens, dtype=np.uint64)
        nw = self.nwords
        got = []
        order = []

        def sink(user, view, first, n, first_occ, n_occ, hint):
            v = view.contents
            order.append((int(first), int(n), int(hint)))
            if n:
                part = dict(first=int(first))
                w = np.ctypeslib.as_array(v.kmer_words, shape=(int(first + n) * nw,))[int(first) * nw:].copy()
                part["words"] = w.reshape(int(n), nw)
                part["cnt"] = np.ctypeslib.as_array(v.cnt, shape=(int(first + n),))[int(first):].copy()
                if self.ext:
                    off = np.ctypeslib.as_array(v.occ_off, shape=(int(first + n),))[int(first):].astype(np.uint64)
                    ends = np.concatenate([off[1:], np.array([first_occ + n_occ], dtype=np.uint64)])
                    part["occ_n"] = ends - off
                    part["pos"] = np.ctypeslib.as_array(v.pos, shape=(int(first_occ + n_occ),))[int(first_occ):].copy() if n_occ else np.zeros(0, np.uint32)
                    part["rid"] = np.ctypeslib.as_array(v.rid, shape=(int(first_occ + n_occ),))[int(first_occ):].copy() if n_occ else np.zeros(0, np.int32)
                got.append(part)
            return 0

        cb = SINK_FN(sink)
        r = Result()
        _check(self.lib.hsk_count_stream(self.handle, packed.ctypes.data, packed.nbytes, readlens.ctypes.data, len(readlens),
                                         readid_base, C.cast(cb, C.c_void_p), None, C.byref(r)))
        got.sort(key=lambda p: p["first"])
        n = int(r.n_kept)
        out = dict(nwords=nw, n_kept=n, n_occ=int(r.n_occ), parts=order,
                   words=np.concatenate([p["words"] for p in got]) if got else np.zeros((0, nw), np.uint64),
                   cnt=np.concatenate([p["cnt"] for p in got]) if got else np.zeros(0, np.uint32),
                   histogram=_arr(r.histogram, self.upper + 1, np.uint64), stats=r.stats.as_dict())
        if self.ext:
            occ_n = np.concatenate([p["occ_n"] for p in got]) if got else np.zeros(0, np.uint64)
            out["occ_off"] = np.concatenate([np.zeros(1, np.uint64), np.cumsum(occ_n, dtype=np.uint64)])
            out["pos"] = np.concatenate([p["pos"] for p in got]) if got else np.zeros(0, np.uint32)
            out["rid"] = np.concatenate([p["rid"] for p in got]) if got else np.zeros(0, np.int32)
        return out

    def count_device(self, d_packed: int, nbytes: int, d_read_off: int, d_read_len: int, nreads: int,
                     readid_base: int = 0) -> DeviceResult:
        """hsk_count_device on device pointers; result stays in HBM."""
        r = DeviceResult()
        _check(self.lib.hsk_count_device(self.handle, d_packed, nbytes, d_read_off, d_read_len, nreads, readid_base,
                                         C.byref(r)))
        return r

    def fetch(self) -> dict:
        r = Result()
        _check(self.lib.hsk_fetch_result(self.handle, C.byref(r)))
        return self._unpack(r)

    def allreduce_histogram(self) -> np.ndarray:
        h = np.zeros(self.upper + 1, dtype=np.uint64)
        _check(self.lib.hsk_allreduce_histogram(self.handle, h.ctypes.data))
        return h

    def fill_entries(self, n: int) -> np.ndarray:
        a = np.zeros((n, self.nwords + 1), dtype=np.uint64)
        _check(self.lib.hsk_fill_entries(self.handle, a.ctypes.data, n))
        return a

    # -- stage-level ---------------------------------------------------------------------------------
    def debug_extract(self, packed: np.ndarray, readlens: np.ndarray, readid_base: int = 0) -> dict:
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        readlens = np.ascontiguousarray(readlens, dtype=np.uint64)
        s = Supermers()
        _check(self.lib.hsk_debug_extract(self.handle, packed.ctypes.data, packed.nbytes, readlens.ctypes.data,
                                          len(readlens), readid_base, C.byref(s)))
        T, S, SW = int(s.n_bins), int(s.n_slots), int(s.slot_words)
        slots = _arr(s.slots, S * SW, np.uint32).reshape(S, SW)
        pw = SW - (2 if self.ext else 0)
        return dict(n_buckets=T, bucket_count=_arr(s.bin_slots, T, np.uint64), bucket_kmers=_arr(s.bin_kmers, T, np.uint64),
                    slot_words=SW, slots=slots, len=(slots[:, pw - 1] & np.uint32(0xFF)).astype(np.int64) if S else np.zeros(0, np.int64),
                    payload=slots[:, :pw].copy() if S else np.zeros((0, pw), np.uint32),
                    ext=((slots[:, SW - 2].astype(np.uint64) << np.uint64(32)) | slots[:, SW - 1].astype(np.uint64)) if (self.ext and S) else None)

    def debug_sort(self, key_ptrs: list[int], tmp_ptrs: list[int], n: int, k: int, val_ptr: int = 0,
                   val_tmp_ptr: int = 0) -> None:
        nw = len(key_ptrs)
        kp = (C.c_void_p * nw)(*key_ptrs)
        tp = (C.c_void_p * nw)(*tmp_ptrs)
        _check(self.lib.hsk_debug_sort(self.handle, kp, tp, val_ptr or None, val_tmp_ptr or None, n, nw, k))
