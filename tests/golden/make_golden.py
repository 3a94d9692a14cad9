"""Generates the golden fixtures in this directory by running the UNMODIFIED reference
(oracle/_ref/*.so, built by oracle/build_ref.sh from /root/reference) on seeded synthetic
reads.  Run from the repo root in the build container: ``python tests/golden/make_golden.py``.

Each fixture holds the input (packed reads + lengths) and the reference's result in canonical
order (k-mers ascending by Kmer::operator<; occurrences ascending by (rid, pos)), plus the
reference's own histogram text (hysortk.cpp:98-136) and the md5 of its sorted output file lines
(hysortk.cpp:149-162).
"""
import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from hysortk_b200 import synth  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = [
    # name, (k, m, l, u, ext), generator
    ("k31_e0_mixed", (31, 17, 2, 50, 0), lambda: synth.sample_mixed(20000, 1600, [30, 31, 97, 150, 200, 263], 0.01, 7)),
    ("k55_e0_mixed", (55, 23, 2, 50, 0), lambda: synth.sample_mixed(20000, 1600, [54, 55, 97, 150, 200, 263], 0.01, 8)),
    ("k31_e1_mixed", (31, 17, 2, 50, 1), lambda: synth.sample_mixed(20000, 1200, [30, 31, 97, 150, 200, 263], 0.01, 9)),
    ("k31_e0_long", (31, 17, 2, 50, 0), lambda: synth.sample_fixed(30000, 12.0, 5000, 0.01, 10)),
    ("k31_e0_lowcomplexity", (31, 17, 2, 50, 0), None),
]


def low_complexity() -> synth.ReadSet:
    """homopolymer / dinucleotide runs: supermers longer than 250 bases (forced split,
    kmerops.cpp:1120), counts above UPPER, N bases (-> A), palindromic context."""
    rng = np.random.Generator(np.random.Philox(11))
    reads = []
    for i in range(60):
        reads.append(synth.ascii_to_codes("A" * 700))
    reads.append(synth.ascii_to_codes("AC" * 400))
    reads.append(synth.ascii_to_codes("GT" * 400))
    reads.append(synth.ascii_to_codes("ACGT" * 100 + "N" * 40 + "ACGT" * 50))
    g = rng.integers(0, 4, 3000, dtype=np.uint8)
    for i in range(40):
        st = int(rng.integers(0, 2000))
        r = g[st:st + 600].copy()
        reads.append(r if i % 2 == 0 else (3 - r[::-1]).astype(np.uint8))
    reads.append(np.zeros(0, dtype=np.uint8))
    reads.append(synth.ascii_to_codes("ACGTTGCA"))
    return synth.pack_reads(reads)


def main() -> None:
    for name, (k, m, l, u, ext), gen in CASES:
        rs = gen() if gen else low_complexity()
        if not po.ref_available(k, m, l, u, ext):
            print("reference build missing for", (k, m, l, u, ext), "- run oracle/build_ref.sh")
            continue
        r = po.ref_kmer_count(rs.packed, rs.readlens, k, m, l, u, ext, want_text=True)
        lines = sorted(r.extra["output_text_raw"].splitlines())
        md5 = hashlib.md5("\n".join(lines).encode()).hexdigest()
        out = dict(packed=rs.packed, readlens=rs.readlens, params=np.array([k, m, l, u, ext], dtype=np.int32),
                   words=r.words, cnt=r.cnt, histogram_text=np.array(r.extra["histogram_text"]),
                   sorted_output_md5=np.array(md5))
        if ext:
            out.update(occ_off=r.occ_off, pos=r.pos, rid=r.rid)
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
        print(name, "reads", rs.nreads, "kept", r.n, "md5", md5)


if __name__ == "__main__":
    main()
