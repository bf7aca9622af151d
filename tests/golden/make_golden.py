#!/usr/bin/env python
"""Golden-vector generator: runs the UNMODIFIED reference (read-only at /root/reference) as the oracle.

This script is the only place that imports the reference.  It cannot run on the GPU box
(/root/reference does not exist there); its outputs are committed under tests/golden/*.npz.

Usage (from the repo root, one process per variant because module names collide between the
two reference directories):

    python tests/golden/make_golden.py --variant unit
    python tests/golden/make_golden.py --variant degree

What is recorded per case (one .npz each):
  * the static multiplex graph exactly as the reference built it (edge lists in G.edges() order,
    num_nodes, max_rank, degree-cost node weights),
  * the greedy rollout of the reference's own GetSol (MultiDismantler_torch.py:759-784): per step the
    action, LMCC rank, running score (f64), MaxCCList entry, numCoveredEdges, len(remove_edge),
    the active-node ids, the f32 Q vector over the active nodes and the aux features,
  * a model-free random-action rollout (MvcEnv.randomAction + stepWithoutReward) pinning the env/LMCC
    alone,
  * the 14-tensor state_dict used.

Four non-invasive shims are injected before import (no edits to the reference tree):
  1. torch_sparse.spmm  -> restatement of rusty1s/pytorch_sparse 0.6.18 spmm (gather * value, scatter-add)
  2. matplotlib stub    (Hyperbolic.py:4-6 imports it at module import)
  3. np.mat = np.asmatrix (PrepareBatchGraph.py:198, removed in NumPy 2)
  4. torch.Tensor.cuda -> identity on this CPU-only container (net_graphsage.py:332,350; mutil_layer_weight.py:276,295)
"""
import argparse
import os
import random
import sys
import types

import numpy as np
import torch

REF = "/root/reference/code"
VARIANT_DIR = {"unit": "MultiDismantler_unit_cost", "degree": "MultiDismantler_degree_cost"}
CKPT = {
    "unit": "models/g0.5_TORCH-Model_GMM_30_50/nrange_30_50_iter_100000.ckpt",
    "degree": "models/nrange_30_50_iter_100000.ckpt",
}
OUT = os.path.dirname(os.path.abspath(__file__))


def install_shims():
    ts = types.ModuleType("torch_sparse")

    def spmm(index, value, m, n, matrix):
        row, col = index[0], index[1]
        out = matrix.index_select(-2, col) * value.unsqueeze(-1)
        res = torch.zeros((m, matrix.size(-1)), dtype=matrix.dtype, device=matrix.device)
        return res.index_add_(0, row, out)

    ts.spmm = spmm
    sys.modules["torch_sparse"] = ts
    mpl = types.ModuleType("matplotlib")
    mpl.use = lambda *a, **k: None
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt
    if not hasattr(np, "mat"):
        np.mat = np.asmatrix
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self


def seed_all(s):
    random.seed(s)
    np.random.seed(s)
    torch.manual_seed(s)


def graph_record(g, variant):
    rec = {
        "num_nodes": np.int64(g.num_nodes),
        "max_rank": np.int64(g.max_rank),
        "edges0": np.array(list(g.edge_list[0]), dtype=np.int32).reshape(-1, 2),
        "edges1": np.array(list(g.edge_list[1]), dtype=np.int32).reshape(-1, 2),
    }
    if variant == "degree":
        for l in range(2):
            rec["weights%d" % l] = np.array([g.weights[l][v] for v in range(g.num_nodes)], dtype=np.float64)
    return rec


class Recorder:
    """Hooks the reference objects; never changes what they compute."""

    def __init__(self, dqn, variant):
        self.dqn = dqn
        self.variant = variant
        self.reset()
        net = dqn.MultiDismantler_net
        orig_fwd = net.test_forward
        orig_setup = dqn.SetuppredAll
        env = dqn.test_env
        orig_step = env.stepWithoutReward
        rec = self

        def fwd(*a, **k):
            out = orig_fwd(*a, **k)
            rec.q.append(out.detach()[:, 0].cpu().numpy().astype(np.float32).copy())
            return out

        def setup(idxes, g_list, covered, remove_edges):
            idx_map_list = orig_setup(idxes, g_list, covered, remove_edges)
            rec.aux.append(dqn.inputs["aux_input"].detach().cpu().numpy().astype(np.float32).copy()[0])
            im = np.array(idx_map_list[0])
            rec.active.append(np.nonzero(im >= 0)[0].astype(np.int32))
            return idx_map_list

        def step(a):
            orig_step(a)
            rec.actions.append(int(a))
            rec.score.append(float(env.score))
            rec.maxcc.append(float(env.MaxCCList[-1]))
            rec.cov_edges.append(list(env.numCoveredEdges))
            rec.rm_edges.append([len(env.remove_edge[0]), len(env.remove_edge[1])])

        net.test_forward = fwd
        dqn.SetuppredAll = setup
        env.stepWithoutReward = step

    def reset(self):
        self.q, self.aux, self.active = [], [], []
        self.actions, self.score, self.maxcc, self.cov_edges, self.rm_edges = [], [], [], [], []

    def dump(self, prefix):
        def ragged(lst, dt):
            off = np.zeros(len(lst) + 1, dtype=np.int64)
            for i, x in enumerate(lst):
                off[i + 1] = off[i] + len(x)
            flat = np.concatenate(lst).astype(dt) if lst else np.zeros(0, dt)
            return flat, off

        qf, qo = ragged(self.q, np.float32)
        af, ao = ragged(self.active, np.int32)
        assert np.array_equal(qo, ao)
        return {
            prefix + "actions": np.array(self.actions, dtype=np.int32),
            prefix + "score": np.array(self.score, dtype=np.float64),
            prefix + "maxcc": np.array(self.maxcc, dtype=np.float64),
            prefix + "cov_edges": np.array(self.cov_edges, dtype=np.int64).reshape(-1, 2),
            prefix + "rm_edges": np.array(self.rm_edges, dtype=np.int64).reshape(-1, 2),
            prefix + "q": qf,
            prefix + "active": af,
            prefix + "off": qo,
            prefix + "aux": np.array(self.aux, dtype=np.float32).reshape(-1, 2, 4),
        }


def run_variant(variant, sizes, seeds_small):
    vdir = os.path.join(REF, VARIANT_DIR[variant])
    sys.path.insert(0, vdir)
    install_shims()
    import MultiDismantler_torch as mdt  # noqa: E402  (does os.chdir(sys.path[0]))
    import graph as refgraph  # noqa: E402

    dqn = mdt.MultiDismantler()
    dqn.LoadModel(os.path.join(vdir, CKPT[variant]))
    sd = {k: v.detach().cpu().numpy().copy() for k, v in dqn.MultiDismantler_net.state_dict().items()}
    np.savez_compressed(os.path.join(OUT, "weights_%s.npz" % variant), **sd)
    rec = Recorder(dqn, variant)

    cases = [(n, 0) for n in sizes] + [(n, s) for (n, s) in seeds_small]
    for (n, seed) in cases:
        seed_all(seed)
        g = refgraph.Graph(n)
        if g.max_rank <= 1:
            print("skip degenerate", n, seed)
            continue
        out = graph_record(g, variant)
        out["seed"] = np.int64(seed)
        # (1) greedy rollout through the reference's own GetSol
        dqn.ClearTestGraphs()
        dqn.InsertGraph(g, is_test=True)
        rec.reset()
        score, sol, cost = dqn.GetSol(0)
        assert [int(a) for a in sol] == rec.actions
        out.update(rec.dump("greedy_"))
        out["greedy_final_score"] = np.float64(score)
        # (2) model-free random rollout (hooks on stepWithoutReward keep recording)
        rec.reset()
        random.seed(1000 + seed)
        env = dqn.test_env
        env.s0(g)
        while not env.isTerminal():
            env.stepWithoutReward(env.randomAction())
        d = rec.dump("random_")
        for k in ("random_actions", "random_score", "random_maxcc", "random_cov_edges", "random_rm_edges"):
            out[k] = d[k]
        name = "%s_n%d_s%d.npz" % (variant, n, seed)
        np.savez_compressed(os.path.join(OUT, name), **out)
        print(name, "edges", len(out["edges0"]), len(out["edges1"]), "max_rank", g.max_rank,
              "greedy steps", len(sol), "AUDC %.10f" % score, "random steps", len(d["random_actions"]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variant", choices=["unit", "degree"], required=True)
    ap.add_argument("--sizes", type=int, nargs="*", default=[30, 50, 100, 200, 1000])
    args = ap.parse_args()
    small = [(40, 1), (45, 2), (60, 3), (150, 4)]
    run_variant(args.variant, args.sizes, small)


if __name__ == "__main__":
    main()
