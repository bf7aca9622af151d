"""Pins oracle/md_oracle.py against golden vectors recorded from the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden_cases, load_golden_graph
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


def test_three_layer_restatement_properties():
    """L = 3 has no reference output (the reference network is hard-wired to two layers); the L-layer restatement is
    checked through properties its definition implies: invariance of Q under a permutation of the layers (fusion and
    the layer gate are symmetric in the layers, every weight is layer-shared), row-sampled head == full head, chunked
    fusion == unchunked fusion, and -- for L = 2 -- the golden vectors above (same code path)."""
    from mdcommunity_b200 import gmm
    N, L = 3000, 3
    layers = gmm.random_sparse_layers(N, L, 4 * N, seed=5)
    w = orc.load_weights(os.path.join(GOLDEN, "weights_unit.npz"))
    envs = []
    for perm in ((0, 1, 2), (2, 0, 1)):
        g = orc.MultiplexGraph(N, [layers[p] for p in perm])
        e = orc.OracleEnv("unit")
        e.s0(g)
        for a in (5, 17, 300):
            e.stepWithoutReward(a)
        envs.append(e)
    assert envs[0].rank_list == envs[1].rank_list and envs[0].score == envs[1].score
    q0, ids0, _ = orc.forward_q(w, [envs[0]], "unit")
    q1, ids1, _ = orc.forward_q(w, [envs[1]], "unit")
    assert np.array_equal(ids0[0], ids1[0])
    assert np.all(np.abs(q0 - q1) <= 2e-5 * np.abs(q0) + 1e-6)
    rows = np.arange(0, len(q0), 7)
    qs, _, _ = orc.forward_q(w, [envs[0]], "unit", head_rows=rows)
    assert np.all(np.abs(qs - q0[rows]) <= 2e-5 * np.abs(q0[rows]) + 1e-6)
    old = orc.FUSE_CHUNK
    try:
        orc.FUSE_CHUNK = 512
        qc, _, _ = orc.forward_q(w, [envs[0]], "unit")
    finally:
        orc.FUSE_CHUNK = old
    assert np.all(np.abs(qc - q0) <= 2e-5 * np.abs(q0) + 1e-6)
    # a batch of two 3-layer graphs == the two graphs alone (batching is per-graph independent)
    qb, idsb, off = orc.forward_q(w, envs, "unit")
    assert np.all(np.abs(qb[:off[1]] - q0) <= 2e-5 * np.abs(q0) + 1e-6)
