"""Seeded synthetic read sets in HySortK's DnaBuffer layout.

Genome = i.i.d. uniform ACGT; reads sampled uniformly, each base substituted with probability
``err`` by a different base, each read reverse-complemented with probability 1/2.  Reads are
produced directly as the packed 2-bit buffer a ``DnaBuffer`` holds (reference
src/dnaseq.cpp:9-31, include/dnabuffer.hpp:14-47: 4 bases per byte, first base in the two most
significant bits, every read starting on a fresh byte, tail bits zero) plus the read lengths,
so large inputs never have to exist as a FASTA file.  ``write_fasta`` emits FASTA + ``.fai`` for
the file-based entry points.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np

_ASCII = np.frombuffer(b"ACGT", dtype=np.uint8)


@dataclass
class ReadSet:
    """Packed reads: ``packed`` (uint8), ``readlens`` (uint64), byte offset of each read."""

    packed: np.ndarray
    readlens: np.ndarray

    @property
    def nreads(self) -> int:
        return int(self.readlens.shape[0])

    @property
    def nbases(self) -> int:
        return int(self.readlens.sum())

    def byte_offsets(self) -> np.ndarray:
        nb = (self.readlens + np.uint64(3)) // np.uint64(4)
        off = np.zeros(self.nreads + 1, dtype=np.uint64)
        np.cumsum(nb, out=off[1:])
        return off

    def num_kmers(self, k: int) -> int:
        ln = self.readlens.astype(np.int64)
        return int(np.maximum(ln - k + 1, 0).sum())

    def codes(self, i: int) -> np.ndarray:
        """2-bit codes of read i (for tests)."""
        off = self.byte_offsets()
        b = self.packed[int(off[i]):int(off[i + 1])]
        c = np.stack([(b >> 6) & 3, (b >> 4) & 3, (b >> 2) & 3, b & 3], axis=1).reshape(-1)
        return c[: int(self.readlens[i])].astype(np.uint8)


def make_genome(length: int, seed: int) -> np.ndarray:
    rng = np.random.Generator(np.random.Philox(seed))
    return rng.integers(0, 4, size=length, dtype=np.uint8)


def pack_codes_matrix(codes: np.ndarray) -> np.ndarray:
    """codes: (n, L) uint8 in 0..3 -> (n, ceil(L/4)) packed bytes."""
    n, L = codes.shape
    pad = (-L) % 4
    if pad:
        codes = np.concatenate([codes, np.zeros((n, pad), dtype=np.uint8)], axis=1)
    c = codes.reshape(n, -1, 4)
    return ((c[:, :, 0] << 6) | (c[:, :, 1] << 4) | (c[:, :, 2] << 2) | c[:, :, 3]).astype(np.uint8)


def pack_reads(reads: list[np.ndarray]) -> ReadSet:
    """Variable-length reads (list of code arrays) -> ReadSet."""
    lens = np.array([len(r) for r in reads], dtype=np.uint64)
    parts = []
    for r in reads:
        if len(r) == 0:
            continue
        parts.append(pack_codes_matrix(np.asarray(r, dtype=np.uint8)[None, :])[0])
    packed = np.concatenate(parts) if parts else np.zeros(0, dtype=np.uint8)
    return ReadSet(np.ascontiguousarray(packed), lens)


def ascii_to_codes(s: str) -> np.ndarray:
    """A/a/N/n->0, C->1, G->2, T->3 (reference include/dnaseq.hpp:138-156)."""
    tab = np.zeros(256, dtype=np.uint8)
    for ch, c in (("A", 0), ("C", 1), ("G", 2), ("T", 3), ("N", 0)):
        tab[ord(ch)] = c
        tab[ord(ch.lower())] = c
    return tab[np.frombuffer(s.encode(), dtype=np.uint8)]


def _mutate_and_flip(reads: np.ndarray, err: float, rng: np.random.Generator) -> np.ndarray:
    n, L = reads.shape
    if err > 0:
        mask = rng.random(size=reads.shape, dtype=np.float32) < err
        shift = rng.integers(1, 4, size=reads.shape, dtype=np.uint8)
        reads = np.where(mask, (reads + shift) & 3, reads).astype(np.uint8)
    flip = rng.random(size=n) < 0.5
    rc = (3 - reads[:, ::-1]).astype(np.uint8)
    return np.where(flip[:, None], rc, reads)


def sample_fixed(genome_len: int, coverage: float, read_len: int, err: float, seed: int,
                 chunk_reads: int = 1 << 16) -> ReadSet:
    """Fixed-length read set: n = genome_len*coverage/read_len reads (vectorised, chunked)."""
    genome = make_genome(genome_len, seed)
    rng = np.random.Generator(np.random.Philox(seed + 1))
    n = int(genome_len * coverage / read_len)
    nb = (read_len + 3) // 4
    packed = np.empty(n * nb, dtype=np.uint8)
    ar = np.arange(read_len, dtype=np.int64)
    for s in range(0, n, chunk_reads):
        e = min(n, s + chunk_reads)
        starts = rng.integers(0, genome_len - read_len + 1, size=e - s, dtype=np.int64)
        reads = genome[starts[:, None] + ar[None, :]]
        reads = _mutate_and_flip(reads, err, rng)
        packed[s * nb:e * nb] = pack_codes_matrix(reads).reshape(-1)
    return ReadSet(packed, np.full(n, read_len, dtype=np.uint64))


def sample_mixed(genome_len: int, nreads: int, lengths: list[int], err: float, seed: int) -> ReadSet:
    """Mixed-length reads (edge coverage: len < K, == K, len%4 in 0..3)."""
    genome = make_genome(genome_len, seed)
    rng = np.random.Generator(np.random.Philox(seed + 1))
    reads = []
    for i in range(nreads):
        L = lengths[i % len(lengths)]
        st = int(rng.integers(0, genome_len - L + 1))
        r = genome[st:st + L][None, :].copy()
        r = _mutate_and_flip(r, err, rng)[0]
        reads.append(r)
    return pack_reads(reads)


def write_fasta(path: str, rs: ReadSet, line_width: int = 0) -> None:
    """FASTA + .fai (name, length, offset, linebases, linewidth); line_width=0 -> single line."""
    off = 0
    with open(path, "wb") as f, open(path + ".fai", "w") as fai:
        for i in range(rs.nreads):
            hdr = f">r{i}\n".encode()
            f.write(hdr)
            off += len(hdr)
            seq = _ASCII[rs.codes(i)].tobytes()
            L = len(seq)
            lw = line_width if line_width > 0 else max(L, 1)
            fai.write(f"r{i}\t{L}\t{off}\t{lw}\t{lw + 1}\n")
            for p in range(0, max(L, 1), lw):
                line = seq[p:p + lw] + b"\n"
                f.write(line)
                off += len(line)


def workload(name: str, seed: int = 42) -> tuple[ReadSet, dict]:
    """Named workloads (BASELINE.json configs; sizes per SURVEY.md §8d)."""
    table = {
        # config[0]: reference's CPU-runnable case, 100 Mbp, 150-bp reads
        "c1_100Mbp_150bp": dict(genome_len=3_340_000, coverage=30.0, read_len=150, err=0.01),
        # config[1]: 30x bacterial-scale, ~150 Mbp, 1% errors (long-read mode, ELBA-style)
        "c2_150Mbp_10kbp": dict(genome_len=5_000_000, coverage=30.0, read_len=10_000, err=0.01),
        "c2_150Mbp_150bp": dict(genome_len=5_000_000, coverage=30.0, read_len=150, err=0.01),
        "tiny_20kbp": dict(genome_len=20_000, coverage=12.0, read_len=150, err=0.01),
        "small_1Mbp": dict(genome_len=200_000, coverage=5.0, read_len=1000, err=0.01),
    }
    p = table[name]
    return sample_fixed(seed=seed, **p), dict(p, name=name, seed=seed)


__all__ = ["ReadSet", "make_genome", "pack_reads", "pack_codes_matrix", "ascii_to_codes", "sample_fixed",
           "sample_mixed", "write_fasta", "workload"]
