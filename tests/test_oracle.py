"""CPU tests of the parity oracle (oracle/oracle.c): against the golden fixtures produced by the
unmodified reference, against the live reference build when present, and internal consistency."""
import hashlib

import numpy as np
import pytest

from hysortk_b200 import synth
from oracle import pyoracle as po


def test_oracle_matches_golden(golden):
    g = golden
    a = po.kmer_count(g["packed"], g["readlens"], g["k"], g["m"], g["lower"], g["upper"], g["ext"], ntasks=5)
    po.assert_equal(a, g["expected"], "oracle vs reference golden")
    assert a.histogram_text() == g["histogram_text"]
    lines = sorted(po.output_text(a).splitlines())
    assert hashlib.md5("\n".join(lines).encode()).hexdigest() == g["sorted_output_md5"]
    # C-side text writers agree with the python ones
    L = po.lib()


def test_supermer_path_equals_direct_definition(golden):
    g = golden
    for ntasks in (1, 5, 23):
        a = po.kmer_count(g["packed"], g["readlens"], g["k"], g["m"], g["lower"], g["upper"], g["ext"], ntasks=ntasks)
        b = po.kmer_count(g["packed"], g["readlens"], g["k"], g["m"], g["lower"], g["upper"], g["ext"],
                          via_supermers=False)
        po.assert_equal(a, b, f"ntasks={ntasks}")
        assert a.total_kmers == b.total_kmers


@pytest.mark.parametrize("k,m,ext", [(31, 17, 0), (55, 23, 0), (31, 17, 1)])
def test_oracle_matches_live_reference(k, m, ext):
    if not po.ref_available(k, m, 2, 50, ext):
        pytest.skip("oracle/_ref not built (needs /root/reference; see oracle/build_ref.sh)")
    rs = synth.sample_fixed(50_000, 8.0, 400, 0.01, seed=100 + k + ext)
    r = po.ref_kmer_count(rs.packed, rs.readlens, k, m, 2, 50, ext, want_text=True)
    a = po.kmer_count(rs.packed, rs.readlens, k, m, 2, 50, ext, ntasks=11)
    po.assert_equal(a, r, "oracle vs live reference")
    assert a.histogram_text() == r.extra["histogram_text"]
    assert a.total_kmers == rs.num_kmers(k)


def test_packing_and_kmer_words():
    """Appendix A of SURVEY.md: base codes, byte packing, k-mer word layout, revcomp, canonical."""
    import ctypes as C
    L = po.lib()
    s = b"ACGTNacgtTTGCA"
    out = np.zeros(4, dtype=np.uint8)
    L.orc_pack_read(s, len(s), out.ctypes.data)
    # A C G T | N(->A) a c g | t T T G | C A 0 0
    assert list(out) == [0b00011011, 0b00000110, 0b11111110, 0b01000000]
    km = (C.c_uint64 * 3)()
    L.orc_kmer_set(out.ctypes.data, 0, 5, km)
    assert km[0] == (0b0001101100 << 54)  # ACGTA left-aligned
    tw = (C.c_uint64 * 3)()
    L.orc_kmer_twin(km, 5, tw)
    buf = C.create_string_buffer(8)
    L.orc_kmer_string(tw, 5, buf)
    assert buf.value == b"TACGT"
    rep = (C.c_uint64 * 3)()
    L.orc_kmer_rep(km, 5, rep)
    L.orc_kmer_string(rep, 5, buf)
    assert buf.value == b"ACGTA"
    # rolling extension == fresh construction, across the 32-base word boundary (K=55)
    rs = synth.sample_fixed(1000, 1.0, 200, 0.0, seed=3)
    mem = rs.packed[:50].copy()
    a = (C.c_uint64 * 3)()
    b = (C.c_uint64 * 3)()
    L.orc_kmer_set(mem.ctypes.data, 0, 55, a)
    codes = rs.codes(0)
    for i in range(1, 100):
        L.orc_kmer_extend(a, 55, int(codes[i + 54]), a)
        L.orc_kmer_set(mem.ctypes.data, i, 55, b)
        assert list(a) == list(b)


def test_murmur_known_answers():
    """MurmurHash3 x64-128 (seed 313) of 8-byte keys: low word, cross-checked with an independent
    python implementation of the published algorithm."""
    L = po.lib()
    M = (1 << 64) - 1

    def rotl(x, r):
        return ((x << r) | (x >> (64 - r))) & M

    def fmix(k):
        k ^= k >> 33; k = (k * 0xff51afd7ed558ccd) & M
        k ^= k >> 33; k = (k * 0xc4ceb9fe1a85ec53) & M
        k ^= k >> 33
        return k

    def mm3(key8: bytes):
        h1 = h2 = 313
        c1, c2 = 0x87c37b91114253d5, 0x4cf5ad432745937f
        k1 = int.from_bytes(key8, "little")
        k1 = (k1 * c1) & M; k1 = rotl(k1, 31); k1 = (k1 * c2) & M; h1 ^= k1
        h1 ^= 8; h2 ^= 8
        h1 = (h1 + h2) & M; h2 = (h2 + h1) & M
        h1 = fmix(h1); h2 = fmix(h2)
        return (h1 + h2) & M

    rng = np.random.default_rng(5)
    for _ in range(50):
        key = rng.integers(0, 256, 8, dtype=np.uint8)
        assert L.orc_murmur3_64(key.ctypes.data, 8) == mm3(key.tobytes())


def test_edge_cases():
    k, m = 31, 17
    # empty input, reads shorter than K, read of exactly K
    rs = synth.pack_reads([])
    a = po.kmer_count(rs.packed, rs.readlens, k, m, 1, 50)
    assert a.n == 0 and a.total_kmers == 0
    r30 = synth.ascii_to_codes("ACGT" * 7 + "AC")
    r31 = synth.ascii_to_codes("ACGT" * 7 + "ACG")
    rs = synth.pack_reads([r30, r31, r30, r31[::-1].copy()])
    a = po.kmer_count(rs.packed, rs.readlens, k, m, 1, 50, ext=1)
    assert a.total_kmers == 2
    # ReadId counts reads shorter than K too (kmerops.cpp:1018)
    assert set(a.rid.tolist()) <= {1, 3}
    # filter edges: count exactly L, U, U+1
    base = synth.make_genome(500, 77)
    for copies, lower, upper, kept in [(3, 3, 3, True), (3, 4, 9, False), (4, 2, 3, False), (1, 1, 1, True)]:
        rs = synth.pack_reads([base[:100]] * copies)
        a = po.kmer_count(rs.packed, rs.readlens, k, m, lower, upper)
        assert (a.n == 70) == kept, (copies, lower, upper, a.n)
        if kept:
            assert np.all(a.cnt == copies)


@pytest.mark.parametrize("k", [3, 17, 31, 32, 33, 55, 64, 65, 77, 95])
def test_device_helpers_on_host_match_oracle(k, tmp_path):
    """kmer_twin / kmer_canonical of hysortk_b200/csrc/common.cuh (the source the kernels compile, run here on the host)
    against the oracle's restatement of Kmer::GetTwin / GetRep (reference include/kmer.hpp:265-303)."""
    import ctypes as C
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    build = os.path.join(root, "tests", "_build")
    os.makedirs(build, exist_ok=True)
    exe = os.path.join(build, "test_common")
    src = os.path.join(root, "tests", "cxx", "test_common.cu")
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(src), os.path.getmtime(os.path.join(root, "hysortk_b200", "csrc", "common.cuh"))):
        subprocess.check_call([os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc"), "-O1", "-std=c++17", src, "-o", exe])
    L = po.lib()
    nw = 1 if k <= 32 else (2 if k <= 64 else 3)
    rng = np.random.default_rng(k)
    lines, kmers = [], []
    for i in range(300):
        codes = rng.integers(0, 4, k)
        if i < 4:   # homopolymers and a palindrome-like pattern
            codes = np.full(k, i % 4)
        w = [0, 0, 0]
        for j, c in enumerate(codes):
            w[j // 32] |= int(c) << (2 * (31 - j % 32))
        kmers.append(w)
        lines.append(f"{k} {w[0]:x} {w[1]:x} {w[2]:x}")
    out = subprocess.run([exe], input="\n".join(lines) + "\n", capture_output=True, text=True, check=True).stdout.splitlines()
    assert len(out) == len(kmers)
    for w, line in zip(kmers, out):
        got = [int(x, 16) for x in line.split()]
        km, tw, rp = (C.c_uint64 * 3)(*w), (C.c_uint64 * 3)(), (C.c_uint64 * 3)()
        L.orc_kmer_twin(km, k, tw)
        L.orc_kmer_rep(km, k, rp)
        assert got[:nw] == list(tw)[:nw] and got[3:3 + nw] == list(rp)[:nw]
