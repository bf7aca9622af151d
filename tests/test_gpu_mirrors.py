"""The reference-shaped host API (MvcEnv / MultiDismantler mirrors) on the device path, against golden vectors."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, assert_sequence_parity, golden_cases

pytestmark = pytest.mark.gpu


def _graph(mdb, z, variant):
    return mdb.Graph_test(z["edges0"], z["edges1"], num_nodes=int(z["num_nodes"]), cost=variant)


@pytest.mark.parametrize("variant", ["unit", "degree"])
def test_mvcenv_mirror_random_trace(variant):
    """MvcEnv.s0 / stepWithoutReward / isTerminal / remove_edge bookkeeping reproduce the reference trace."""
    import mdcommunity_b200 as mdb
    for p in golden_cases(variant, max_n=200):
        z = np.load(p)
        g = _graph(mdb, z, variant)
        env = mdb.MvcEnv(50)
        env.s0(g)
        assert g.max_rank == int(z["max_rank"])
        for t, a in enumerate(z["random_actions"]):
            assert not env.isTerminal()
            env.stepWithoutReward(int(a))
            assert env.score == z["random_score"][t]
            assert env.MaxCCList[-1] == z["random_maxcc"][t]
            assert list(env.numCoveredEdges) == list(z["random_cov_edges"][t])
            assert [len(s) for s in env.remove_edge] == list(z["random_rm_edges"][t])
        assert env.isTerminal()
        assert env.action_list == [int(a) for a in z["random_actions"]]
        with pytest.raises(AssertionError):
            env.stepWithoutReward(int(z["random_actions"][0]))  # already covered (mvc_env.py:76-77)


def test_mvcenv_step_rewards_and_random_action():
    import random
    import mdcommunity_b200 as mdb
    from oracle import md_oracle as orc
    z = np.load(os.path.join(GOLDEN, "unit_n100_s0.npz"))
    g = _graph(mdb, z, "unit")
    og = orc.MultiplexGraph(int(z["num_nodes"]), [z["edges0"], z["edges1"]])
    env, oenv = mdb.MvcEnv(50), orc.OracleEnv("unit")
    env.s0(g)
    oenv.s0(og)
    random.seed(3)
    while not env.isTerminal():
        a = env.randomAction()
        active, _, _ = oenv.status()
        assert env.avail_list == list(np.nonzero(active)[0])
        r = env.step(a)
        assert r == oenv.stepWithoutReward(a)
        assert env.state_seq[-1] == env.action_list[:-1]
    assert oenv.isTerminal()
    assert len(env.reward_seq) == len(env.act_seq) == len(env.state_seq_edges)


@pytest.mark.parametrize("variant", ["unit", "degree"])
def test_driver_getsol_and_predict(variant):
    import mdcommunity_b200 as mdb
    from oracle import md_oracle as orc
    dqn = mdb.MultiDismantler(cost=variant)
    dqn.LoadModel(os.path.join(GOLDEN, "weights_%s.npz" % variant))
    for gid, p in enumerate(golden_cases(variant, max_n=200)):
        z = np.load(p)
        g = _graph(mdb, z, variant)
        dqn.InsertGraph(g, is_test=True)
        score, sol, cost = dqn.GetSol(gid)
        ref = [int(a) for a in z["greedy_actions"]]
        assert_sequence_parity(sol, score, z, (variant, gid))
        # Predict on the reference's states (covered list + remove_edge sets rebuilt by the MvcEnv mirror)
        env = mdb.MvcEnv(50)
        env.s0(g)
        off = z["greedy_off"]
        for t in range(min(3, len(ref))):
            pred = dqn.Predict([g], [env.action_list], [env.remove_edge])[0]
            ids = z["greedy_active"][off[t]:off[t + 1]]
            qg = z["greedy_q"][off[t]:off[t + 1]]
            assert np.all(pred[np.setdiff1d(np.arange(g.num_nodes), ids)] == -(2147483647 / 2))
            assert np.all(np.abs(pred[ids] - qg) <= 1e-4 * np.abs(qg) + 1e-6)
            env.stepWithoutReward(ref[t])
    dqn.ClearTestGraphs()


def test_getsolution_topk_matches_oracle():
    """GetSolution with step > 1 (testReal's stepRatio mode): same removal batches as the oracle's top-k rule."""
    import mdcommunity_b200 as mdb
    from oracle import md_oracle as orc
    z = np.load(os.path.join(GOLDEN, "unit_n200_s0.npz"))
    g = _graph(mdb, z, "unit")
    dqn = mdb.MultiDismantler(cost="unit")
    dqn.LoadModel(os.path.join(GOLDEN, "weights_unit.npz"))
    dqn.InsertGraph(g, is_test=True)
    sol, score, maxcc = dqn.GetSolution(0, step=4)
    ow = orc.load_weights(os.path.join(GOLDEN, "weights_unit.npz"))
    og = orc.MultiplexGraph(int(z["num_nodes"]), [z["edges0"], z["edges1"]])
    r = orc.rollout(ow, og, "unit", forced_actions=sol)  # LMCC / AUDC of OUR sequence: bit-exact
    assert r["score"] == score and r["maxcc"] == maxcc
    free = orc.rollout(ow, og, "unit", step=4, record=True)
    same = 0
    while same < min(len(sol), len(free["sol"])) and sol[same] == free["sol"][same]:
        same += 1
    assert same >= 4  # at least the first batch is identical; later ones may differ on rounding-level ties


def test_real_file_formats(tmp_path):
    import mdcommunity_b200 as mdb
    z = np.load(os.path.join(GOLDEN, "unit_n60_s3.npz"))
    path = tmp_path / "toy_multiplex.edges"
    with open(path, "w") as f:
        for l, key in ((1, "edges0"), (2, "edges1")):
            for u, v in z[key]:
                f.write("%d %d %d\n" % (l, u + 1, v + 1))
            f.write("%d 3 3\n" % l)  # self loop: dropped like the reference's read_multiplex
    dqn = mdb.MultiDismantler(cost="unit")
    dqn.LoadModel(os.path.join(GOLDEN, "weights_unit.npz"))
    sol, secs, score = dqn.EvaluateRealData(str(path), str(tmp_path), 0.0, int(z["num_nodes"]), (1, 2))
    assert_sequence_parity(sol, score, z, "EvaluateRealData")
    out = tmp_path / "StepRatio_0.0000"
    lines = open(out / "Soluion_toy_multiplex_12.txt").read().split()
    assert [int(x) for x in lines] == sol
    lm = open(out / "NormalizedLMCC_toy_multiplex_12.txt").read().split()
    assert len(lm) == int(z["num_nodes"]) + 2 and float(lm[0]) == 1.0


def test_two_envs_on_one_graph_do_not_share_state():
    """The reference keeps all episode state per env (covered_set / remove_edge / score): several MvcEnv objects may hold
    the same graph (env_list + TrainSet.Sample() with replacement; test_env next to a training env).  Interleaved
    episodes on one graph object must each reproduce the golden trace."""
    import mdcommunity_b200 as mdb
    z = np.load(os.path.join(GOLDEN, "unit_n100_s0.npz"))
    g = _graph(mdb, z, "unit")
    e1, e2 = mdb.MvcEnv(50), mdb.MvcEnv(50)
    e1.s0(g)
    e2.s0(g)
    assert e1._ro is not e2._ro and e1._ro.batch is e2._ro.batch  # device CSR shared, episode state per env
    acts = [int(a) for a in z["random_actions"]]
    greedy = [int(a) for a in z["greedy_actions"]]
    for t in range(max(len(acts), len(greedy))):
        if t < len(acts):
            e1.stepWithoutReward(acts[t])
            assert e1.score == z["random_score"][t] and e1.MaxCCList[-1] == z["random_maxcc"][t]
        if t < len(greedy):
            e2.stepWithoutReward(greedy[t])
            assert e2.MaxCCList[-1] == z["greedy_maxcc"][t]
    assert e1.isTerminal() and e2.isTerminal()
    assert e2.score == float(z["greedy_final_score"])
    e1.s0(g)  # re-use: a fresh episode on the same env
    assert e1.score == 0.0 and not e1.isTerminal() and e2.isTerminal()


def _dense(n, edges):
    a = np.zeros((n, n), np.float64)
    a[edges[:, 0], edges[:, 1]] = 1.0
    a[edges[:, 1], edges[:, 0]] = 1.0
    return a


@pytest.mark.parametrize("variant", ["unit", "degree"])
def test_evaluate_npy_sets_and_test_score(variant, tmp_path):
    """Evaluate with the reference's signature on .npy dense-adjacency test sets (U/MultiDismantler_torch.py:563-600,
    file layout :575-576) and Test (:738-755): aggregates equal the per-graph GetSol results, every rollout replays
    bit-exactly in the oracle, Test = AUDC + remaining-node tail with the reference's expression."""
    import mdcommunity_b200 as mdb
    from mdcommunity_b200 import graph as mg
    from oracle import md_oracle as orc
    cases = golden_cases(variant, max_n=100)[:3]
    d = tmp_path / "data_g" / "syn_64"
    os.makedirs(d)
    for i, p in enumerate(cases):
        z = np.load(p)
        n = int(z["num_nodes"])
        np.save(d / ("adj1_%d.npy" % i), _dense(n, z["edges0"]))
        np.save(d / ("adj2_%d.npy" % i), _dense(n, z["edges1"]))
    dqn = mdb.MultiDismantler(cost=variant)
    wpath = os.path.join(GOLDEN, "weights_%s.npz" % variant)
    out = dqn.Evaluate("unused", "64", "data_g", wpath, data_root=str(tmp_path), n_test=len(cases))
    score_mean, score_std, time_mean, time_std, cost_mean = out
    ow = orc.load_weights(wpath)
    scores, costs = [], []
    for i in range(len(cases)):
        g = mg.load_npy_pair(str(d / ("adj1_%d.npy" % i)), str(d / ("adj2_%d.npy" % i)), cost=variant)
        dqn.InsertGraph(g, is_test=True)
        s, sol, c = dqn.GetSol(i)
        og = orc.MultiplexGraph(g.num_nodes, g.edge_list, weights=g.weights if variant == "degree" else None)
        r = orc.rollout(ow, og, variant, forced_actions=sol)
        assert r["score"] == s and r["sol"] == sol
        # free-running oracle on the same (re-ordered) edge lists: equal up to the first near-tie
        free = orc.rollout(ow, og, variant, record=True)
        same = 0
        while same < min(len(sol), len(free["sol"])) and sol[same] == free["sol"][same]:
            same += 1
        if not (same == len(sol) == len(free["sol"])):
            ids, q = list(free["active"][same]), free["q"][same]
            gap = abs(float(q[ids.index(free["sol"][same])]) - float(q[ids.index(sol[same])]))
            assert gap <= 1e-4 * abs(float(q.max())) + 1e-6
        scores.append(s)
        costs.append(c)
        # Test(): greedy with the first-max rule (== GetSol's device rule) + the tail of the remaining nodes (:752-753)
        t = dqn.Test(i)
        remain = sorted(set(range(g.num_nodes)) ^ set(sol))
        if variant == "unit":
            assert t == s + len(remain) / (g.max_rank * g.num_nodes)
            assert c == len(sol) / g.num_nodes
        else:
            assert t >= s
    assert score_mean == np.mean(scores) and score_std == np.std(scores) and cost_mean == np.mean(costs)
    assert time_mean > 0 and time_std >= 0
    dqn.ClearTestGraphs()
    # list-of-graphs convenience form gives the same aggregates
    gs = [mg.load_npy_pair(str(d / ("adj1_%d.npy" % i)), str(d / ("adj2_%d.npy" % i)), cost=variant) for i in range(len(cases))]
    out2 = dqn.Evaluate(gs)
    assert out2[0] == score_mean and out2[4] == cost_mean


def test_sharded_getsolbatch_two_gpus(tmp_path):
    """Product-level sharded rollout (MultiDismantler.GetSolBatch under torchrun, NCCL): 2 ranks, 9 graphs round-robin,
    one all_gather of per-graph records; every rank must hold the single-GPU results bit for bit."""
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    script = os.path.join(os.path.dirname(GOLDEN), "sharded_getsolbatch_worker.py")
    out = tmp_path / "out"
    os.makedirs(out)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29731", script, str(out)]
    subprocess.run(cmd, check=True, timeout=600)
    import mdcommunity_b200 as mdb
    from mdcommunity_b200 import gmm
    graphs = [gmm.gmm_graph(150 + 10 * i, 40 + i, cost="unit") for i in range(9)]
    dqn = mdb.MultiDismantler(cost="unit")
    dqn.LoadModel(os.path.join(GOLDEN, "weights_unit.npz"))
    score, sols, res = dqn.GetSolBatch(graphs)
    for r in range(2):
        z = np.load(out / ("rank%d.npz" % r))
        assert np.array_equal(z["score"], score)
        assert np.array_equal(z["n_steps"], res["n_steps"])
        T = min(z["actions"].shape[1], res["actions"].shape[1])
        assert np.array_equal(z["actions"][:, :T], res["actions"][:, :T])
        assert np.array_equal(z["max_rank"], [g.max_rank for g in graphs])


def test_graphbatch_upload_streams_a_new_batch_into_resident_buffers():
    """GraphBatch.upload: another batch of the same shape is copied into the resident device buffers (no allocation; the
    rollout object and its captured CUDA graph stay bound).  The same graphs in reversed order have the same shape and a
    different content: the rollout after the upload must equal the reversed results bit for bit, twice (eager, then the
    CUDA-graph replay), and a batch of another shape is refused."""
    import mdcommunity_b200 as mdb
    from mdcommunity_b200 import gmm
    cost = "degree"
    graphs = [gmm.gmm_graph(90 + 7 * i, 300 + i, cost=cost) for i in range(6)]
    w = mdb.QWeights(dict(np.load(os.path.join(GOLDEN, "weights_%s.npz" % cost))), torch.device("cuda:0"))
    batch = mdb.GraphBatch(graphs, cost=cost)
    ro = mdb.BatchRollout(batch, w)
    r1 = ro.run()
    fresh = [mdb.Graph_test(g.edge_list[0], g.edge_list[1], num_nodes=g.num_nodes, cost=cost) for g in graphs[::-1]]
    host2 = mdb.GraphBatch.prepare_host(fresh, cost)  # LMCC of the intact graphs not known yet: reset() fills it in
    for rep in range(3):
        batch.upload(host2)
        ro.reset()
        r2 = ro.run()
        for k in ("score", "n_steps", "actions"):
            assert np.array_equal(r2[k], r1[k][::-1]), (rep, k)
        for i in range(len(graphs)):  # (trace columns beyond n_steps keep whatever an earlier rollout left there)
            n = int(r2["n_steps"][i])
            j = len(graphs) - 1 - i
            assert np.array_equal(r2["ranks"][i, :n], r1["ranks"][j, :n]) and np.array_equal(r2["maxcc"][i, :n], r1["maxcc"][j, :n]), (rep, i)
        assert [g.max_rank for g in fresh] == [g.max_rank for g in graphs[::-1]]
    other = mdb.GraphBatch.prepare_host([gmm.gmm_graph(90, 1, cost=cost) for _ in range(6)], cost)
    with pytest.raises(ValueError):
        batch.upload(other)
