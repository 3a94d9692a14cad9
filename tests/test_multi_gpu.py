"""Multi-GPU parity (needs >= 2 GPUs; skipped on a 1-GPU box): runs tools/mgpu_check.py under
torchrun on all visible GPUs (capped at 8)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_union_over_ranks_equals_oracle():
    import torch
    n = min(torch.cuda.device_count(), 8)
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.join(ROOT, "tools", "mgpu_check.py")],
                       capture_output=True, text=True, timeout=900)
    assert "MGPU_CHECK_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.gpu
def test_standalone_cli_as_mpi_job(tmp_path):
    """`./hysortk reads.fa outdir` (make standalone) as a 2-rank MPI job, one rank per GPU, started by hsk_mpirun with the
    bundled MPI stand-in: read_dna_buffer gives every rank its share of the FASTA, kmer_count exchanges the supermers, every
    rank writes <rank>.out.  The union of the output files and the histogram equal the reference's (golden fixture)."""
    import hashlib
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import load_golden
    from test_cxx_api import BUILD, make
    from hysortk_b200 import synth
    g = load_golden("k31_e0_mixed")
    _, binp = make(g["k"], g["m"], g["lower"], g["upper"], g["ext"], log=1)
    run = os.path.join(BUILD, "hsk_mpirun")
    subprocess.check_call(["g++", "-O2", "-std=c++17", os.path.join(ROOT, "hysortk_b200", "shim", "hsk_mpirun.cpp"), "-o", run, "-lrt"])
    rs = synth.ReadSet(g["packed"], g["readlens"])
    keep = [i for i in range(rs.nreads) if rs.readlens[i] > 0]   # FASTA cannot hold empty records portably
    rs = synth.pack_reads([rs.codes(i) for i in keep])
    fasta = str(tmp_path / "reads.fa")
    synth.write_fasta(fasta, rs, 70)
    outdir = tmp_path / "out"
    outdir.mkdir()
    r = subprocess.run([run, "-n", "2", "-t", "4", binp, fasta, str(outdir)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert g["histogram_text"] in r.stdout
    lines = sorted(open(outdir / "0.out").read().splitlines() + open(outdir / "1.out").read().splitlines())
    assert len(open(outdir / "0.out").read()) > 0 and len(open(outdir / "1.out").read()) > 0
    assert hashlib.md5("\n".join(lines).encode()).hexdigest() == g["sorted_output_md5"]
