"""MvcEnv with the reference's call surface, backed by the device-resident environment kernels.

Mirror of /root/reference/code/MultiDismantler_unit_cost/mvc_env.py:8-162 (degree-cost reward:
/root/reference/code/MultiDismantler_degree_cost/mvc_env.py:68-133): same methods (s0, step, stepWithoutReward,
randomAction, isTerminal, getReward, getMaxConnectedNodesNum), same readable fields (graph, action_list, covered_set,
remove_edge, numCoveredEdges, score, MaxCCList, state_seq, state_seq_edges, act_seq, reward_seq, sum_rewards) and the
same assertions on misuse.  One env = a batch of one graph; many-graph rollouts should use BatchRollout directly.
"""
import random

import numpy as np
import torch

from .graph import GraphBatch
from .rollout import BatchRollout


class MvcEnv:
    def __init__(self, norm=None, device=None, track_remove_edge=True):
        self.norm = norm
        self.device = device
        self.track_remove_edge = track_remove_edge
        self.graph = None
        self.numCoveredEdges = [0, 0]
        self.state_seq, self.act_seq, self.action_list = [], [], []
        self.reward_seq, self.sum_rewards = [], []
        self.covered_set = set()
        self.avail_list = []
        self.remove_edge = [set(), set()]
        self.state_seq_edges = []
        self.MaxCCList = [1]
        self.score = 0.0
        self._ro = None

    # ------------------------------------------------------------------------------------------------
    def s0(self, _g):
        """Reset (mvc_env.py:31-52): clears the episode and runs the initial mutual-component prune on the GPU."""
        self.graph = _g
        # The immutable device CSR is cached on the graph object; the mutable episode state (labels, counters, score,
        # traces) belongs to THIS env: several envs may hold the same graph at once (env_list + TrainSet.Sample() with
        # replacement, test_env next to a training env), exactly like the reference's per-env covered_set / remove_edge.
        batch = getattr(_g, "_mdb_batch", None)
        if batch is None or (self.device is not None and torch.device(self.device) != batch.device):
            batch = GraphBatch([_g], device=self.device, cost=_g.cost)
            _g._mdb_batch = batch
        if self._ro is not None and self._ro.batch is batch:
            self._ro.reset()
        else:
            self._ro = BatchRollout(batch, None)
        cached = self._ro
        L = _g.num_layers
        self.covered_set.clear()
        self.action_list.clear()
        self.numCoveredEdges = [0] * L
        self.state_seq.clear()
        self.act_seq.clear()
        self.reward_seq.clear()
        self.sum_rewards.clear()
        self.remove_edge = [set() for _ in range(L)]
        self.state_seq_edges.clear()
        self.MaxCCList = [1]
        self.score = 0.0
        self._act = torch.zeros(1, 1, dtype=torch.int32, device=cached.batch.device)
        self._alive = None
        if self.track_remove_edge:
            self._alive = [np.ones(len(e), bool) for e in _g.edge_list]
            self._sync_remove_edge(None)

    def _labels(self):
        return self._ro.label[:self.graph.num_nodes].cpu().numpy()

    def _sync_remove_edge(self, a):
        """Rebuild the reference's remove_edge sets (Mcc.deledge, Mcc.py:3-10): edges that stopped being alive in this
        step without touching the removed node were cut by the mutual-component prune."""
        lab = self._labels()
        for l, e in enumerate(self.graph.edge_list):
            u, v = e[:, 0], e[:, 1]
            alive = (lab[u] >= 0) & (lab[u] == lab[v])
            gone = self._alive[l] & ~alive
            if a is not None:
                gone &= (u != a) & (v != a)
            for x, y in e[gone]:
                self.remove_edge[l].add((int(x), int(y)))
                self.remove_edge[l].add((int(y), int(x)))
            self._alive[l] = alive

    def _advance(self, a):
        assert self.graph
        a = int(a)
        assert a not in self.covered_set
        self._act[0, 0] = a
        self._ro.step(self._act, 0)
        self.covered_set.add(a)
        self.action_list.append(a)
        ro = self._ro
        rank = int(ro.rank[0])
        self.numCoveredEdges = [int(x) for x in ro.cov_edges[:, 0].cpu().numpy()]
        if self.track_remove_edge:
            self._sync_remove_edge(a)
        return rank

    def _reward(self, rank, a):
        g = self.graph
        if g.cost == "degree":  # D/mvc_env.py:127-133
            w0, w1 = g.weights[0], g.weights[1]
            tw0, tw1 = float(sum(w0.tolist())), float(sum(w1.tolist()))
            return -float(rank) / (g.max_rank) * (float(w0[a]) / tw0 + float(w1[a]) / tw1) / 2.0, float(rank) / g.max_rank
        return -float(rank) / (g.max_rank * float(g.num_nodes)), None  # U/mvc_env.py:133-137

    def step(self, a):
        """mvc_env.py:53-72: records the transition for the replay memory and returns the reward."""
        assert self.graph
        assert a not in self.covered_set
        self.state_seq.append(self.action_list.copy())
        self.state_seq_edges.append([s.copy() for s in self.remove_edge])
        self.act_seq.append(a)
        rank = self._advance(a)
        r_t, _ = self._reward(rank, int(a))
        self.reward_seq.append(r_t)
        self.sum_rewards.append(r_t)
        return r_t

    def stepWithoutReward(self, a):
        """mvc_env.py:74-87.  score / MaxCCList come from the device accumulators (bit-exact f64)."""
        rank = self._advance(a)
        self.score = float(self._ro.score[0])
        t = int(self._ro.n_steps[0]) - 1
        self.MaxCCList.append(float(self._ro.trace_maxcc[0, t]))

    def randomAction(self):
        """mvc_env.py:89-101: uniform choice among the un-covered nodes with an alive edge in every layer
        (== the active list, ascending), using Python's `random` like the reference."""
        assert self.graph
        n = int(self._ro.n_act[0])
        self.avail_list = self._ro.act_list[:n].cpu().numpy().tolist()
        assert self.avail_list
        return random.choice(self.avail_list)

    def isTerminal(self):
        assert self.graph
        return bool(int(self._ro.done[0]))

    def getMaxConnectedNodesNum(self, a=None):
        return float(int(self._ro.rank[0]))

    def getReward(self, a):
        r, _ = self._reward(int(self._ro.rank[0]), int(a))
        return r
