"""Oracle restatement of the heuristic baselines (HDA / CI) vs golden vectors recorded from the reference scripts
(tests/golden/make_golden_baselines.py -> tests/golden/baselines_unit.npz).  CPU only."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden_graph


def _cases():
    z = np.load(os.path.join(GOLDEN, "baselines_unit.npz"))
    return z, sorted({k.split("/")[0] for k in z.files})


@pytest.mark.parametrize("kind", ["hda", "ci"])
def test_heuristic_oracle_matches_reference(kind):
    from oracle import md_oracle as orc
    z, cases = _cases()
    assert len(cases) >= 6
    for case in cases:
        _, _, g = load_golden_graph(os.path.join(GOLDEN, case + ".npz"))
        r = orc.heuristic_rollout(g, kind, "max")
        assert r["sol"] == [int(a) for a in z[case + "/%s_sol" % kind]], case
        # ND_mcc of the scripts = [1] + LMCC_t / LMCC_0 ; MaxCCList of the env = [1] + (LMCC_t / (LMCC_0 N)) * N
        ref = z[case + "/%s_mcc" % kind]
        assert len(ref) == len(r["maxcc"])
        assert [int(round(x * g.max_rank)) for x in ref[1:]] == r["ranks"], case
        if kind == "hda":
            assert r["score"] == float(z[case + "/hda_score"]), case  # same f64 operation order


def test_heuristic_add_variants_terminate():
    from oracle import md_oracle as orc
    z, cases = _cases()
    _, _, g = load_golden_graph(os.path.join(GOLDEN, cases[0] + ".npz"))
    for kind in ("hda", "ci"):
        r = orc.heuristic_rollout(g, kind, "add")
        assert r["ranks"][-1] == 1 and len(set(r["sol"])) == len(r["sol"])
