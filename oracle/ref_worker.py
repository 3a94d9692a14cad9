"""TEST INFRASTRUCTURE — one MPI rank of a multi-rank run of the UNMODIFIED reference (oracle/_ref), started by
pyoracle.ref_kmer_count_ranks with HSK_MPI_SIZE / HSK_MPI_RANK / HSK_MPI_SESSION in its environment (the bundled
multi-process MPI stand-in, hysortk_b200/shim/mpi.h).

    python ref_worker.py in.npz out.npz K M L U EXT repeats want_result

Prints one JSON line {"seconds": [...]} (the reference's kmer_count wall time of every repeat on this rank)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pyoracle as po  # noqa: E402


def main():
    inp, outp = sys.argv[1], sys.argv[2]
    k, m, lower, upper, ext, repeats, want = [int(x) for x in sys.argv[3:10]]
    z = np.load(inp)
    packed, lens = z["packed"], z["readlens"]
    # the reference logs to stdout: keep the JSON line apart
    real = os.dup(1)
    os.dup2(2, 1)
    secs, c = [], None
    for _ in range(repeats):
        c = po.ref_kmer_count(packed, lens, k, m, lower, upper, ext)
        secs.append(c.seconds)
    if want:
        # raw per-rank result (canonicalize() only sorted it): the union is re-sorted by the caller
        kw = dict(words=c.words, cnt=c.cnt)
        if ext:
            kw.update(occ_off=c.occ_off, pos=c.pos, rid=c.rid)
        np.savez(outp, **kw)
    os.dup2(real, 1)
    sys.stdout = os.fdopen(real, "w")
    print(json.dumps({"seconds": secs}), flush=True)


if __name__ == "__main__":
    main()
