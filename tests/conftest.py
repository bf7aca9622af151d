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


Q_RTOL, Q_ATOL = 1e-4, 1e-6  # north_star: Q within 1e-4 relative; 1e-6 absolute floor (|q| ~ 1e-3)


def assert_sequence_parity(sol, score, z, ctx=""):
    """The stated rule for a free-running removal sequence against the reference's golden greedy trace `z`:
    identical sequence => identical AUDC (bit-exact); otherwise the FIRST differing pick must be a near-tie in the
    reference's own Q vector: |q_ref_pick - q_our_pick| <= 1e-4 |q_max| + 1e-6.  Never passes vacuously."""
    ref = [int(a) for a in z["greedy_actions"]]
    sol = [int(a) for a in sol]
    same = 0
    while same < min(len(sol), len(ref)) and sol[same] == ref[same]:
        same += 1
    if same == len(ref) == len(sol):
        assert score == float(z["greedy_final_score"]), (ctx, score, float(z["greedy_final_score"]))
        return same
    assert same < len(ref) and same < len(sol), (ctx, "one sequence is a strict prefix of the other", same, len(sol), len(ref))
    off = z["greedy_off"]
    qg = z["greedy_q"][off[same]:off[same + 1]]
    ids = list(z["greedy_active"][off[same]:off[same + 1]])
    assert sol[same] in ids, (ctx, "picked an inactive node", same, sol[same])
    gap = abs(float(qg[ids.index(ref[same])]) - float(qg[ids.index(sol[same])]))
    assert gap <= Q_RTOL * abs(float(qg.max())) + Q_ATOL, (ctx, same, gap)
    return same
