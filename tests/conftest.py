import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_cases(variant, max_n=None):
    out = []
    for p in sorted(glob.glob(os.path.join(GOLDEN, "%s_n*_s*.npz" % variant))):
        n = int(os.path.basename(p).split("_n")[1].split("_")[0])
        if max_n is None or n <= max_n:
            out.append(p)
    return out


def load_golden_graph(path):
    from oracle import md_oracle as orc
    z = np.load(path)
    variant = os.path.basename(path).split("_")[0]
    weights = [z["weights0"], z["weights1"]] if variant == "degree" else None
    g = orc.MultiplexGraph(int(z["num_nodes"]), [z["edges0"], z["edges1"]], weights=weights,
                           max_rank=None)
    return z, variant, g


@pytest.fixture(scope="session")
def oracle_weights():
    from oracle import md_oracle as orc
    return {v: orc.load_weights(os.path.join(GOLDEN, "weights_%s.npz" % v)) for v in ("unit", "degree")}
