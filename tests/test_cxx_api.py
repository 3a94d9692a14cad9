"""The C++ drop-in layer: public value types (host only), the Makefile artefacts, and — on the GPU —
the standalone CLI end to end against the reference's golden output."""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np
import pytest

from hysortk_b200 import synth
from oracle import pyoracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tests", "_build")


def make(k, m, l, u, ext, target="standalone", log=0):
    tag = f"k{k}_m{m}_l{l}_u{u}_e{ext}_log{log}"
    obj = os.path.join(BUILD, "obj_" + tag)
    binp = os.path.join(BUILD, "hysortk_" + tag)
    subprocess.check_call(["make", "-s", "-j8", target, f"K={k}", f"M={m}", f"L={l}", f"U={u}", f"EXT={ext}", f"LOG={log}",
                           f"OBJ={obj}", f"BIN={binp}", "CUOBJ=" + os.path.join(BUILD, "cuda_obj")], cwd=ROOT,
                          stdout=subprocess.DEVNULL)
    return obj, binp


@pytest.mark.parametrize("k,ext", [(31, 0), (55, 0), (31, 1), (70, 0), (32, 0)])
def test_value_types_match_oracle(k, ext):
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, f"test_types_k{k}_e{ext}")
    subprocess.check_call(["g++", "-O1", "-std=c++17", f"-DKMER_SIZE={k}", "-DMINIMIZER_SIZE=17", "-DLOWER_KMER_FREQ=2",
                           "-DUPPER_KMER_FREQ=50", f"-DEXTENSION={ext}", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cxx", "test_types.cpp"),
                           os.path.join(ROOT, "hysortk_b200", "cxx", "dnaseq.cpp"),
                           os.path.join(ROOT, "hysortk_b200", "cxx", "dnabuffer.cpp"),
                           os.path.join(ROOT, "hysortk_b200", "cxx", "hashfuncs.cpp"), "-o", exe])
    rng = np.random.default_rng(k)
    reads = ["".join("ACGT"[c] for c in rng.integers(0, 4, n)) for n in (k, k + 1, 150, 2 * k + 3)]
    reads.append("ACGTN" * 30)
    out = subprocess.run([exe] + reads, capture_output=True, text=True, check=True).stdout.splitlines()
    L = po.lib()
    nw = 1 if k <= 32 else (2 if k <= 64 else 3)
    it = iter(out)
    hdr = next(it).split()
    assert int(hdr[1]) == len(reads)
    assert int(hdr[3]) == sum((len(r) + 3) // 4 for r in reads) == int(hdr[5])
    for r in reads:
        line = next(it).split()
        assert line[1] == r.replace("N", "A") and int(line[2]) == (len(r) + 3) // 4
        packed = np.zeros((len(r) + 3) // 4, dtype=np.uint8)
        L.orc_pack_read(r.encode(), len(r), packed.ctypes.data)
        km, tw, rp = (C.c_uint64 * 3)(), (C.c_uint64 * 3)(), (C.c_uint64 * 3)()
        buf = C.create_string_buffer(128)
        for p in range(len(r) - k + 1):
            got = next(it).split()
            L.orc_kmer_set(packed.ctypes.data, p, k, km)
            L.orc_kmer_twin(km, k, tw)
            L.orc_kmer_rep(km, k, rp)
            exp = []
            for x in (km, tw, rp):
                L.orc_kmer_string(x, k, buf)
                exp.append(buf.value.decode())
            assert got[:3] == exp
            assert int(got[3]) == L.orc_murmur3_64(rp, 8 * nw)
    sizes = next(it).split()
    assert int(sizes[1]) == 8 * nw
    if not ext:
        assert int(sizes[2]) == 8 * nw + 8 and int(sizes[3]) == 8 * nw
    else:
        assert int(sizes[2]) == 8 * nw + 8 + 48 and int(sizes[3]) == 8 * nw + 8   # two std::vector + (pos, rid)


def test_makefile_artefacts():
    obj, binp = make(31, 17, 2, 50, 0)
    assert os.path.exists(os.path.join(obj, "libhysortk.o")) and os.path.exists(binp)
    syms = subprocess.run(["nm", "-C", os.path.join(obj, "libhysortk.o")], capture_output=True, text=True).stdout
    for s in ["hysortk::kmer_count(", "hysortk::read_dna_buffer(", "hysortk::print_kmer_histogram(",
              "hysortk::write_output_file(", "hsk_count", "hsk_create"]:
        assert s in syms, s
    r = subprocess.run(["make", "K=17", "M=17"], cwd=ROOT, capture_output=True, text=True)
    assert r.returncode != 0 and "must be less than" in (r.stderr + r.stdout)   # reference Makefile:50-52


@pytest.mark.parametrize("line_width", [0, 60, 7])
def test_read_dna_buffer_bytes(line_width, tmp_path):
    """read_dna_buffer (host threads encode the records in place) yields exactly the packed bytes of the reads,
    whatever the FASTA line width; no GPU involved."""
    obj, _ = make(31, 17, 2, 50, 0, target="lib")
    exe = os.path.join(BUILD, "test_read")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-fopenmp", "-DKMER_SIZE=31", "-DMINIMIZER_SIZE=17", "-DLOWER_KMER_FREQ=2",
                           "-DUPPER_KMER_FREQ=50", "-DEXTENSION=0", "-DLOG_LEVEL=0", "-I" + os.path.join(ROOT, "include"),
                           "-I" + os.path.join(ROOT, "hysortk_b200", "shim"), os.path.join(ROOT, "tests", "cxx", "test_read.cpp"),
                           os.path.join(obj, "libhysortk.o"), "-L" + os.path.join(cuda, "lib64"), "-lcudart", "-ldl", "-lpthread",
                           "-o", exe])
    rng = np.random.default_rng(line_width + 1)
    lens = [1, 2, 3, 4, 5, 30, 31, 59, 60, 61, 120, 121, 1000, 4097] + [int(x) for x in rng.integers(1, 700, 300)]
    reads = [rng.integers(0, 4, n).astype(np.uint8) for n in lens]
    rs = synth.pack_reads(reads)
    fasta = str(tmp_path / "reads.fa")
    synth.write_fasta(fasta, rs, line_width)
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(cuda, "lib64") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    out = subprocess.run([exe, fasta], capture_output=True, text=True, check=True, env=env).stdout.splitlines()
    n, bufsize = (int(x) for x in out[0].split())
    assert n == len(reads) and bufsize == rs.packed.nbytes
    off = rs.byte_offsets()
    for i, line in enumerate(out[1:]):
        parts = line.split()
        assert int(parts[0]) == lens[i]
        got = bytes.fromhex(parts[1]) if len(parts) > 1 else b""
        assert got == rs.packed[int(off[i]):int(off[i + 1])].tobytes(), i


@pytest.mark.parametrize("k", [31, 55])
def test_write_output_file_lines(k, tmp_path):
    """write_output_file (blocks formatted by the host threads) writes "<k-mer>\\t<count>" per entry in list order; no GPU."""
    obj, _ = make(k, 17, 2, 50, 0, target="lib")
    exe = os.path.join(BUILD, f"test_write_k{k}")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-fopenmp", f"-DKMER_SIZE={k}", "-DMINIMIZER_SIZE=17", "-DLOWER_KMER_FREQ=2",
                           "-DUPPER_KMER_FREQ=50", "-DEXTENSION=0", "-DLOG_LEVEL=0", "-I" + os.path.join(ROOT, "include"),
                           "-I" + os.path.join(ROOT, "hysortk_b200", "shim"), os.path.join(ROOT, "tests", "cxx", "test_write.cpp"),
                           os.path.join(obj, "libhysortk.o"), "-L" + os.path.join(cuda, "lib64"), "-lcudart", "-ldl", "-lpthread",
                           "-o", exe])
    rng = np.random.default_rng(k)
    n = 3000
    kmers = ["".join("ACGT"[c] for c in rng.integers(0, 4, k)) for _ in range(n)]
    cnts = [int(x) for x in rng.integers(1, 70000, n)]
    cnts[:4] = [1, 9, 10, 65535]
    inp = "".join(f"{a} {c}\n" for a, c in zip(kmers, cnts))
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(cuda, "lib64") + ":" + os.environ.get("LD_LIBRARY_PATH", ""), OMP_NUM_THREADS="5")
    subprocess.run([exe, str(tmp_path)], input=inp, text=True, check=True, env=env)
    got = open(tmp_path / "0.out").read()
    assert got == "".join(f"{a}\t{c}\n" for a, c in zip(kmers, cnts))


@pytest.mark.gpu
@pytest.mark.parametrize("name,line_width", [("k31_e0_mixed", 0), ("k55_e0_mixed", 80), ("k31_e0_lowcomplexity", 60)])
def test_standalone_cli_matches_reference_output(name, line_width, tmp_path):
    from conftest import load_golden
    g = load_golden(name)
    _, binp = make(g["k"], g["m"], g["lower"], g["upper"], g["ext"], log=1)
    rs = synth.ReadSet(g["packed"], g["readlens"])
    keep = [i for i in range(rs.nreads) if rs.readlens[i] > 0]   # FASTA cannot hold empty records portably
    if len(keep) != rs.nreads:
        rs = synth.pack_reads([rs.codes(i) for i in keep])
    fasta = str(tmp_path / "reads.fa")
    synth.write_fasta(fasta, rs, line_width)
    outdir = tmp_path / "out"
    outdir.mkdir()
    r = subprocess.run([binp, fasta, str(outdir)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert g["histogram_text"] in r.stdout
    assert "Overall kmer counting (Excluding I/O)" in r.stdout
    lines = sorted(open(outdir / "0.out").read().splitlines())
    assert hashlib.md5("\n".join(lines).encode()).hexdigest() == g["sorted_output_md5"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["k31_e0_mixed", "k55_e0_mixed", "k31_e1_mixed"])
def test_cxx_kmer_count_matches_reference_golden(name, tmp_path, monkeypatch):
    """hysortk::kmer_count itself (C++ API over the engine; C entry points hysortk_b200/cxx/bench_api.cpp): pageable
    DnaBuffer in, std::vector<KmerListEntryS> out — entries incl. the per-entry pos / rid vectors of EXTENSION == 1, the
    histogram text and the output file equal the unmodified reference's (tests/golden)."""
    from conftest import load_golden
    from hysortk_b200 import cxxapi
    g = load_golden(name)
    a = cxxapi.kmer_count(g["packed"], g["readlens"], g["k"], g["m"], g["lower"], g["upper"], g["ext"], want_text_dir=str(tmp_path))
    got = po.canonicalize(g["k"], a["words"], a["cnt"], a.get("occ_off"), a.get("pos"), a.get("rid"))
    po.assert_equal(got, g["expected"], "C++ API vs reference golden")
    assert a["n"] == g["expected"].n
    assert open(tmp_path / "hist.txt").read() == g["histogram_text"]
    lines = sorted(open(tmp_path / "0.out").read().splitlines())
    assert hashlib.md5("\n".join(lines).encode()).hexdigest() == g["sorted_output_md5"]
    # a second call on the same engine, and a larger input (several result parts, list grown from the hint)
    rs = synth.sample_fixed(500_000, 12.0, 2500, 0.01, seed=6)
    exp = po.kmer_count(rs.packed, rs.readlens, g["k"], g["m"], g["lower"], g["upper"], g["ext"], via_supermers=False)
    b = cxxapi.kmer_count(rs.packed, rs.readlens, g["k"], g["m"], g["lower"], g["upper"], g["ext"])
    po.assert_equal(po.canonicalize(g["k"], b["words"], b["cnt"], b.get("occ_off"), b.get("pos"), b.get("rid")), exp, "C++ API vs oracle")
    # the engine's estimate of the list size falls short (forced): the parts that do not fit are left out and the list is
    # rebuilt from the complete result after the call
    monkeypatch.setenv("HSK_TEST_LIST_HINT_DIV", "3")
    c = cxxapi.kmer_count(rs.packed, rs.readlens, g["k"], g["m"], g["lower"], g["upper"], g["ext"])
    po.assert_equal(po.canonicalize(g["k"], c["words"], c["cnt"], c.get("occ_off"), c.get("pos"), c.get("rid")), exp, "C++ API, list rebuilt")
    assert np.array_equal(c["words"], b["words"]) and np.array_equal(c["cnt"], b["cnt"])
