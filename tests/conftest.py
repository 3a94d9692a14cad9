import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ["k31_e0_mixed", "k55_e0_mixed", "k31_e1_mixed", "k31_e0_long", "k31_e0_lowcomplexity"]


def load_golden(name):
    import numpy as np
    from oracle import pyoracle as po
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    k, m, l, u, ext = [int(x) for x in z["params"]]
    nw = 1 if k <= 32 else (2 if k <= 64 else 3)
    exp = po.Counts(k, nw, z["words"], z["cnt"], z["occ_off"] if ext else None, z["pos"] if ext else None,
                    z["rid"] if ext else None)
    return dict(packed=z["packed"], readlens=z["readlens"], k=k, m=m, lower=l, upper=u, ext=ext, expected=exp,
                histogram_text=str(z["histogram_text"]), sorted_output_md5=str(z["sorted_output_md5"]))


@pytest.fixture(params=GOLDEN_CASES)
def golden(request):
    return load_golden(request.param)
