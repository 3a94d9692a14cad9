"""The bundled multi-process MPI stand-in (hysortk_b200/shim/mpi.h + hsk_mpirun): every collective the kmer_count path
uses, on several ranks of this host, and the UNMODIFIED reference run as a 2- and 3-rank MPI job through it —
BASELINE.json configs[0] is "2 MPI ranks on CPU" — against the golden fixtures (which come from a 1-rank run)."""
import os
import subprocess

import pytest

from oracle import pyoracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tests", "_build")


@pytest.mark.parametrize("nranks,slot_mb", [(1, 4), (2, 4), (5, 4), (3, 1)])
def test_collectives(nranks, slot_mb):
    os.makedirs(BUILD, exist_ok=True)
    exe, run = os.path.join(BUILD, "test_mpi_shim"), os.path.join(BUILD, "hsk_mpirun")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I" + os.path.join(ROOT, "hysortk_b200", "shim"),
                           os.path.join(ROOT, "tests", "cxx", "test_mpi_shim.cpp"), "-o", exe, "-lrt"])
    subprocess.check_call(["g++", "-O2", "-std=c++17", os.path.join(ROOT, "hysortk_b200", "shim", "hsk_mpirun.cpp"), "-o", run, "-lrt"])
    env = dict(os.environ, HSK_MPI_SLOT_MB=str(slot_mb), HSK_MPI_TIMEOUT_S="60")
    r = subprocess.run([run, "-n", str(nranks), exe], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and f"shim ok n={nranks}" in r.stdout, r.stdout + r.stderr
    # a failing rank takes the job down instead of leaving the others waiting
    r = subprocess.run([run, "-n", "3", exe, "abort"], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode != 0


@pytest.mark.parametrize("name", ["k31_e0_mixed", "k55_e0_mixed", "k31_e1_mixed"])
@pytest.mark.parametrize("nranks", [2, 3])
def test_reference_as_mpi_job_matches_golden(name, nranks):
    from conftest import load_golden
    g = load_golden(name)
    if not po.ref_available(g["k"], g["m"], g["lower"], g["upper"], g["ext"]):
        pytest.skip("oracle/_ref not built")
    c, secs = po.ref_kmer_count_ranks(g["packed"], g["readlens"], g["k"], g["m"], g["lower"], g["upper"], g["ext"], nranks=nranks,
                                      threads_per_rank=2)
    po.assert_equal(c, g["expected"], f"reference, {nranks} MPI ranks")
    assert len(secs) == 1 and secs[0] > 0


def test_reference_multi_round_exchange_matches_oracle():
    """A read set large enough for many 80 KB exchange rounds per rank pair (the reference's MPI_Ialltoall / MPI_Wait loop
    with a barrier in between, kmerops.cpp:814-968): the reference as a 4-rank job over the MPI stand-in equals the C
    oracle, which has no notion of ranks."""
    from hysortk_b200 import synth
    if not po.ref_available(31, 17, 2, 50, 0):
        pytest.skip("oracle/_ref not built")
    rs = synth.sample_fixed(300_000, 15.0, 1000, 0.01, seed=12)
    exp = po.kmer_count(rs.packed, rs.readlens, 31, 17, 2, 50, 0, via_supermers=False)
    c, secs = po.ref_kmer_count_ranks(rs.packed, rs.readlens, 31, 17, 2, 50, 0, nranks=4, threads_per_rank=2)
    po.assert_equal(c, exp, "reference, 4 MPI ranks, vs oracle")
    # ... and the extension fields: global ReadIds come from the MPI_Exscan of the read counts (kmerops.cpp:65-70)
    if po.ref_available(31, 17, 2, 50, 1):
        exp1 = po.kmer_count(rs.packed, rs.readlens, 31, 17, 2, 50, 1, via_supermers=False)
        c1, _ = po.ref_kmer_count_ranks(rs.packed, rs.readlens, 31, 17, 2, 50, 1, nranks=3, threads_per_rank=2)
        po.assert_equal(c1, exp1, "reference EXT, 3 MPI ranks, vs oracle")
