"""Pins oracle/md_oracle.py against golden vectors recorded from the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest

from conftest import golden_cases, load_golden_graph
from oracle import md_oracle as orc

ALL = golden_cases("unit") + golden_cases("degree")


@pytest.mark.parametrize("path", ALL, ids=[os.path.basename(p) for p in ALL])
def test_env_random_rollout_bit_exact(path):
    """Model-free trace: LMCC, AUDC score, MaxCCList, covered / pruned edge counters are bit-exact."""
    z, variant, g = load_golden_graph(path)
    assert g.max_rank == int(z["max_rank"])  # Graph.ori_rank (U/graph.py:80-84)
    env = orc.OracleEnv(variant)
    env.s0(g)
    for t, a in enumerate(z["random_actions"]):
        assert not env.isTerminal()
        env.stepWithoutReward(int(a))
        assert env.score == z["random_score"][t]
        assert env.MaxCCList[-1] == z["random_maxcc"][t]
        assert list(env.numCoveredEdges) == list(z["random_cov_edges"][t])
        assert [2 * env.num_removed_edges(l) for l in range(2)] == list(z["random_rm_edges"][t])
    assert env.isTerminal()


@pytest.mark.parametrize("path", ALL, ids=[os.path.basename(p) for p in ALL])
def test_greedy_rollout_teacher_forced(path, oracle_weights):
    """Feed the reference's actions; per step compare active set, aux (exact), Q (rtol 2e-5), score (exact)."""
    z, variant, g = load_golden_graph(path)
    w = oracle_weights[variant]
    env = orc.OracleEnv(variant)
    env.s0(g)
    off = z["greedy_off"]
    agree = 0
    for t, a in enumerate(z["greedy_actions"]):
        assert not env.isTerminal()
        active, _, aux = env.status()
        ids = np.nonzero(active)[0]
        assert np.array_equal(ids, z["greedy_active"][off[t]:off[t + 1]])
        assert np.array_equal(aux, z["greedy_aux"][t])
        q = orc.predict(w, [env], variant)[0]
        qg = z["greedy_q"][off[t]:off[t + 1]]
        np.testing.assert_allclose(q[ids].astype(np.float32), qg, rtol=2e-5, atol=1e-6)
        # the oracle's pick equals the reference's unless the top-two gap is a rounding-level tie
        pick = orc.arg_first_max(q)
        if pick == int(a):
            agree += 1
        else:
            top = np.sort(qg)[::-1]
            assert abs(q[pick] - q[int(a)]) <= 1e-6 * max(1.0, abs(top[0])), (t, pick, int(a))
        env.stepWithoutReward(int(a))
        assert env.score == z["greedy_score"][t]
        assert env.MaxCCList[-1] == z["greedy_maxcc"][t]
        assert list(env.numCoveredEdges) == list(z["greedy_cov_edges"][t])
    assert env.isTerminal()
    assert env.score == float(z["greedy_final_score"])
    assert agree >= 0.5 * len(z["greedy_actions"])


def test_free_running_rollout_n100(oracle_weights):
    """Free-running oracle rollout reproduces the reference AUDC on the survey's N=100 case."""
    import conftest
    path = os.path.join(conftest.GOLDEN, "unit_n100_s0.npz")
    z, variant, g = load_golden_graph(path)
    r = orc.rollout(oracle_weights[variant], g, variant)
    assert abs(r["score"] - float(z["greedy_final_score"])) < 1e-12
    assert r["sol"][:10] == list(z["greedy_actions"][:10])
