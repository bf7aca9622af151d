"""Heuristic baselines (HDA / CI) on the device vs the reference scripts' golden output and the pinned oracle."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden_cases, load_golden_graph

pytestmark = pytest.mark.gpu


def _graphs(max_n=None):
    import mdcommunity_b200 as mdb
    zs, gs, ogs = [], [], []
    for p in golden_cases("unit", max_n):
        z, _, og = load_golden_graph(p)
        zs.append((os.path.basename(p)[:-4], z))
        gs.append(mdb.Graph_test(z["edges0"], z["edges1"], num_nodes=int(z["num_nodes"]), cost="unit"))
        ogs.append(og)
    return zs, gs, ogs


@pytest.mark.parametrize("kind", ["hda", "ci"])
def test_heuristic_matches_reference_golden(kind):
    """Removal sequence, normalised LMCC list and (HDA) the f64 score of the reference scripts, bit-exact, on a ragged
    batch (all golden graphs up to N = 200 at once)."""
    from mdcommunity_b200 import baselines
    ref = np.load(os.path.join(GOLDEN, "baselines_unit.npz"))
    zs, gs, _ = _graphs(200)
    out = baselines.critical_number_batch(gs, kind, "max")
    for (case, z), g, (nd, sol, score) in zip(zs, gs, out):
        assert sol == [int(a) for a in ref[case + "/%s_sol" % kind]], case
        assert np.array_equal(np.array(nd), ref[case + "/%s_mcc" % kind]), case
        if kind == "hda":
            assert score == float(ref[case + "/hda_score"]), case


@pytest.mark.parametrize("kind,combine", [("hda", "max"), ("hda", "add"), ("ci", "max"), ("ci", "add")])
def test_heuristic_matches_oracle(kind, combine):
    """All four variants, including the N = 1000 graph, against the oracle restatement: sequence, LMCC trace, AUDC."""
    from oracle import md_oracle as orc
    from mdcommunity_b200 import baselines
    zs, gs, ogs = _graphs()
    out = baselines.critical_number_batch(gs, kind, combine)
    for (case, z), og, (nd, sol, score) in zip(zs, ogs, out):
        r = orc.heuristic_rollout(og, kind, combine)
        assert sol == r["sol"], (case, kind, combine)
        assert [int(round(x * og.max_rank)) for x in nd[1:]] == r["ranks"], case
        assert score == r["score"], case


def test_critical_number_single_graph_signature():
    from mdcommunity_b200 import baselines
    ref = np.load(os.path.join(GOLDEN, "baselines_unit.npz"))
    z = np.load(os.path.join(GOLDEN, "unit_n50_s0.npz"))
    nd, sol, score = baselines.critical_number(z["edges0"], z["edges1"], int(z["num_nodes"]), int(z["max_rank"]))
    assert sol == [int(a) for a in ref["unit_n50_s0/hda_sol"]] and score == float(ref["unit_n50_s0/hda_score"])
