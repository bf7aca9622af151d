"""MultiDismantler driver with the reference's call surface for the rollout path.

Mirror of /root/reference/code/MultiDismantler_unit_cost/MultiDismantler_torch.py (class MultiDismantler :66-843) and
its degree-cost twin, restricted to the greedy-rollout path: LoadModel :791-797, InsertGraph / ClearTestGraphs
:166-181, Predict :263-302, GetSol :759-784, GetSolution :711-736, Test :738-755, argMax :799-807, Evaluate :563-600
(on caller-supplied graphs: the reference's ../../data .npy sets are not part of its repository), read_multiplex
:602-635 and the Soluion_/NormalizedLMCC_ writers of EvaluateRealData :645-709.

Extra entry point: GetSolBatch(graphs) runs many independent rollouts in lock step on the device -- the reference can
only do one graph at a time (g_list has length 1, :760-768).
"""
import os
import time

import numpy as np
import torch

from . import graph as mgraph
from .env import MvcEnv
from .graph import GraphBatch
from .rollout import BatchRollout, QWeights

inf = 2147483647 / 2  # MultiDismantler_torch.py:60
NUM_MAX = 50


class MultiDismantler:
    def __init__(self, cost="unit", device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("mdcommunity_b200.MultiDismantler needs a CUDA device (sm_100a); no CPU fallback")
        assert cost in ("unit", "degree")
        self.cost = cost
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.TestSet = mgraph.GSet()
        self.ngraph_test = 0
        self.test_env = MvcEnv(NUM_MAX, device=self.device)
        self.weights = None
        self.state_dict = None

    # ---- model ------------------------------------------------------------------------------------------
    def LoadModel(self, model_path):
        """Accepts the reference's .ckpt (torch state_dict, 14 keys) unchanged, or an .npz with the same keys."""
        if isinstance(model_path, dict):
            sd = model_path
        elif str(model_path).endswith(".npz"):
            z = np.load(model_path)
            sd = {k: z[k] for k in z.files}
        else:
            sd = torch.load(model_path, map_location="cpu")
        self.state_dict = sd
        self.weights = QWeights(sd, self.device)

    # ---- graph pool ----------------------------------------------------------------------------------------
    def InsertGraph(self, g, is_test=True):
        assert is_test, "only the test pool exists on the rollout path"
        t = self.ngraph_test
        self.ngraph_test += 1
        self.TestSet.InsertGraph(t, g)

    def ClearTestGraphs(self):
        self.ngraph_test = 0
        self.TestSet.Clear()

    # ---- scoring ---------------------------------------------------------------------------------------------
    def _rollout_for(self, env):
        ro = env._ro
        ro.w = self.weights
        return ro

    def Predict(self, g_list, covered, remove_edges=None, isSnapSnot=False):
        """Q values scattered to [N] float64 with -inf for inactive / covered nodes (:263-302), one array per graph.
        `covered[i]` is the list of removed nodes; `remove_edges[i]` the reference's per-layer sets of pruned directed
        edges (only used, as in get_status_info, to split dead edges into pruned vs covered for the aux features).
        Unlike the reference (which resets `pred` inside the 64-graph chunk loop, :265/:288) all graphs are returned."""
        assert self.weights is not None, "LoadModel first"
        batch = GraphBatch(g_list, device=self.device, cost=self.cost)
        ro = BatchRollout(batch, self.weights)
        B, L, S = batch.B, batch.L, batch.S
        mask = np.zeros(B * S, np.uint8)
        cov = np.zeros((L, B), np.int32)
        for i, g in enumerate(g_list):
            c = np.asarray(list(covered[i]), dtype=np.int64)
            mask[i * S + c] = 1
            cm = np.zeros(g.num_nodes, bool)
            cm[c] = True
            for l in range(L):
                e = g.edge_list[l]
                touched = cm[e[:, 0]] | cm[e[:, 1]]
                if remove_edges is not None and remove_edges[i] is not None and len(remove_edges[i][l]):
                    rm = remove_edges[i][l]
                    pruned = np.fromiter(((int(x), int(y)) in rm for x, y in e), bool, len(e))
                    touched &= ~pruned
                cov[l, i] = int(touched.sum())  # `counter` of get_status_info (PrepareBatchGraph.py:45-52)
        ro.cov_edges.copy_(torch.from_numpy(cov))
        dmask = torch.from_numpy(mask).to(self.device)
        ro.set_state(dmask)
        ro.forward()
        q = ro.q.cpu().numpy()
        act = ro.act_list.cpu().numpy()
        nact = ro.n_act.cpu().numpy()
        out = []
        for i, g in enumerate(g_list):
            p = np.full(g.num_nodes, -inf, np.float64)
            ids = act[i * S:i * S + nact[i]]
            p[ids - i * S] = q[ids]
            out.append(p)
        return out

    def PredictWithCurrentQNet(self, g_list, covered, remove_edges=None):
        return self.Predict(g_list, covered, remove_edges, False)

    @staticmethod
    def argMax(scores):
        """First index of the maximum (:799-807)."""
        return int(np.argmax(np.asarray(scores)))

    # ---- rollouts ----------------------------------------------------------------------------------------------
    def _greedy(self, gid, step):
        assert self.weights is not None, "LoadModel first"
        env = self.test_env
        env.track_remove_edge = False
        env.s0(self.TestSet.Get(gid))
        ro = self._rollout_for(env)
        res = ro.run(step=step)
        n = int(res["n_steps"][0])
        sol = [int(a) for a in res["actions"][0, :n]]
        env.action_list[:] = sol
        env.covered_set = set(sol)
        env.score = float(res["score"][0])
        env.MaxCCList = [1] + [float(x) for x in res["maxcc"][0, :n]]
        env.numCoveredEdges = [int(x) for x in ro.cov_edges[:, 0].cpu().numpy()]
        return env, sol

    def GetSol(self, gid, step=1):
        """(:759-784) -> (AUDC score, removal sequence, cost).  cost = |sol| / N for unit cost, the summed node cost
        for degree cost (D/MultiDismantler_torch.py:734-768)."""
        env, sol = self._greedy(gid, step)
        g = env.graph
        if self.cost == "unit":
            return env.score, sol, len(sol) / g.num_nodes
        tw0, tw1 = float(sum(g.weights[0].tolist())), float(sum(g.weights[1].tolist()))
        total_cost_value = 0
        for v in sol:
            total_cost_value += (g.weights[0][v] / tw0 + g.weights[1][v] / tw1) / 2.0
        return env.score, sol, total_cost_value

    def GetSolution(self, gid, step=1):
        """(:711-736) top-`step` nodes per forward -> (solution, score, MaxCCList)."""
        env, sol = self._greedy(gid, step)
        return sol, env.score, env.MaxCCList

    def Test(self, gid):
        """(:738-755) greedy rollout with the first-max rule; score plus the tail of remaining nodes."""
        env, sol = self._greedy(gid, 1)
        g = env.graph
        remain = sorted(set(range(g.num_nodes)) ^ set(sol))
        if self.cost == "unit":
            return env.score + len(remain) / (g.max_rank * g.num_nodes)
        tw0, tw1 = float(sum(g.weights[0].tolist())), float(sum(g.weights[1].tolist()))
        remain_score = 0
        for v in remain:
            remain_score += 1 / g.max_rank * (g.weights[0][v] / tw0 + g.weights[1][v] / tw1) / 2
        return env.score + remain_score

    def GetSolBatch(self, graphs, step=1, distributed=None):
        """Rollouts of many independent graphs: returns (scores f64 [B], list of removal sequences, results dict).

        distributed: None = shard when torch.distributed is initialised with more than one rank.  Sharded mode is the
        product-level multi-GPU entry point of the path: graph i goes to rank i mod W (sharding.shard_indices), every rank
        rolls out its own graphs with no data-path collective, and ONE all_gather of fixed-width per-graph records (AUDC
        f64 bit-exact, n_steps, removal sequence, LMCC trace) returns the complete result on every rank."""
        import torch.distributed as dist
        from . import sharding
        assert self.weights is not None, "LoadModel first"
        graphs = list(graphs)
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        if distributed is None:
            distributed = world > 1
        if not distributed or world == 1:
            batch = GraphBatch(graphs, device=self.device, cost=self.cost)
            ro = BatchRollout(batch, self.weights)
            res = ro.run(step=step)
        else:
            rank = dist.get_rank()
            mine = sharding.shard_indices(len(graphs), rank, world)
            width = max(g.num_nodes for g in graphs)
            per_rank_max = (len(graphs) + world - 1) // world
            if mine:
                batch = GraphBatch([graphs[i] for i in mine], device=self.device, cost=self.cost)
                ro = BatchRollout(batch, self.weights, max_steps=width)
                r = ro.run(step=step)
                for j, i in enumerate(mine):
                    graphs[i].max_rank = batch.graphs[j].max_rank
                rec = sharding.pack_records(r["score"], r["n_steps"], r["actions"], r["ranks"], width)
                rec = torch.cat([rec, torch.tensor([[float(batch.graphs[j].max_rank)] for j in range(len(mine))],
                                                   dtype=torch.float64)], 1)
            else:
                rec = torch.zeros(0, 3 + 2 * width, dtype=torch.float64)
            full = sharding.gather_records(rec, len(graphs), per_rank_max, device=self.device)
            score, n_steps, actions, ranks = sharding.unpack_records(full[:, :2 + 2 * width], width)
            max_rank = full[:, 2 + 2 * width].cpu().numpy().astype(np.int64)
            for i, g in enumerate(graphs):
                g.max_rank = int(max_rank[i])
            res = dict(score=score, n_steps=n_steps, actions=actions, ranks=ranks, max_rank=max_rank)
        B = len(graphs)
        sols = [[int(a) for a in res["actions"][i, :int(res["n_steps"][i])]] for i in range(B)]
        return res["score"], sols, res

    def Evaluate(self, data_test, data_test_name=None, data_type=None, model_file=None, data_root="../../data/synthetic",
                 n_test=20):
        """(:563-600) mean/std of AUDC, of the per-graph solve time and the mean cost over the synthetic test set
        `{data_root}/{data_type}/syn_{data_test_name}/adj{1,2}_{i}.npy`, i < n_test (the reference hard-codes
        data_root = "../../data/synthetic" and n_test = 20, 2 in its smoke mode; `data_test` is unused there as well).
        Timed region per graph = GetSol, as in the reference (:588-590).
        Convenience: `data_test` may also be a list of Graph_test objects (then no files are read)."""
        if model_file is not None:
            self.LoadModel(model_file)
        assert self.weights is not None, "LoadModel first (the reference's findModel scans a training directory)"
        if isinstance(data_test, (list, tuple)) and (not data_test or isinstance(data_test[0], mgraph.Graph_test)):
            graphs = list(data_test)
        else:
            if data_test_name is None or data_type is None:
                raise ValueError("Evaluate(data_test, data_test_name, data_type, model_file) as in the reference")
            d = os.path.join(data_root, str(data_type), "syn_%s" % data_test_name)
            graphs = [mgraph.load_npy_pair(os.path.join(d, "adj1_%s.npy" % i), os.path.join(d, "adj2_%s.npy" % i), cost=self.cost)
                      for i in range(n_test)]
        self.ClearTestGraphs()
        scores, times, costs = [], [], []
        for i, g in enumerate(graphs):
            self.InsertGraph(g, is_test=True)
            t1 = time.time()
            val, sol, cost_value = self.GetSol(i)
            torch.cuda.synchronize(self.device)
            t2 = time.time()
            scores.append(val)
            times.append(t2 - t1)
            costs.append(cost_value)
        self.ClearTestGraphs()
        return np.mean(scores), np.std(scores), np.mean(times), np.std(times), np.mean(costs)

    # ---- testReal file formats ------------------------------------------------------------------------------------
    @staticmethod
    def read_multiplex(path, N):
        """.edges text (`layer u v`, 1-based, layers in ascending blocks; self loops dropped) -> per-layer edge arrays
        (:602-635).  Duplicate edges collapse like in an nx.Graph."""
        layers = {}
        order = []
        with open(path, "r") as f:
            for line in f:
                el = line.strip(" \n").split(" ")
                if len(el) < 3:
                    continue
                lid, a, b = int(el[0]), int(el[1]) - 1, int(el[2]) - 1
                if lid not in layers:
                    layers[lid] = {}
                    order.append(lid)
                if a == b:
                    continue
                key = (a, b) if (b, a) not in layers[lid] else (b, a)
                layers[lid].setdefault(key, None)
        return [np.array(list(layers[l].keys()), dtype=np.int64).reshape(-1, 2) for l in order]

    def EvaluateRealData(self, edges_path, save_dir, stepRatio, num_nodes, layers, model_file=None):
        """(:645-709) -> (solution, seconds, score); writes Soluion_*.txt and NormalizedLMCC_*.txt like the reference."""
        if model_file is not None:
            self.LoadModel(model_file)
        test_name = os.path.basename(edges_path)
        save_dir_local = os.path.join(save_dir, "StepRatio_%.4f" % stepRatio)
        os.makedirs(save_dir_local, exist_ok=True)
        stem = test_name.split(".")[0]
        f1 = "%s/%s_%s_%s%s.%s" % (save_dir_local, "Soluion", stem, layers[0], layers[1], "txt")
        f2 = "%s/%s_%s_%s%s.%s" % (save_dir_local, "NormalizedLMCC", stem, layers[0], layers[1], "txt")
        lay = self.read_multiplex(edges_path, num_nodes)
        g = mgraph.Graph_test(lay[layers[0] - 1], lay[layers[1] - 1], num_nodes=num_nodes, cost=self.cost)
        step = max(int(stepRatio * g.num_nodes), 1) if stepRatio > 0 else 1
        self.InsertGraph(g, is_test=True)
        t1 = time.time()
        solution, score, MaxCCList = self.GetSolution(self.ngraph_test - 1, step)
        torch.cuda.synchronize(self.device)
        t2 = time.time()
        with open(f1, "w") as f:
            for a in solution:
                f.write("%d\n" % a)
        with open(f2, "w") as f:
            for j in range(g.num_nodes):
                f.write("%.8f\n" % (float(MaxCCList[j]) if j < len(MaxCCList) else 1 / g.max_rank))
            f.write("%.8f\n" % score)
            f.write("%.8f\n" % 0.0)
        self.ClearTestGraphs()
        return solution, t2 - t1, score
