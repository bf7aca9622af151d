#!/usr/bin/env python
"""Golden vectors for one DQN training step: runs the UNMODIFIED reference
(MultiDismantler_unit_cost) on a replay-style batch built from the committed golden graphs and records
q_pred, the fused node embeddings, the loss of calc_loss and the gradient of every parameter, plus the parameters
after one Adam step.

Batch: for each golden graph with N <= 60, two transitions (state = first t greedy removals, action = the
(t+1)-th greedy removal) with fixed targets.  Calls, exactly as `fit` does (MultiDismantler_torch.py:381-408):
SetupTrain -> train_forward -> calc_loss -> backward -> optimizer.step.

Only this script imports the reference (same four shims as make_golden.py); output: tests/golden/train_unit.npz.

    python tests/golden/make_golden_train.py
"""
import glob
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (shims, seeds)


def main():
    variant = "unit"
    vdir = os.path.join(mg.REF, mg.VARIANT_DIR[variant])
    sys.path.insert(0, vdir)
    mg.install_shims()
    import MultiDismantler_torch as mdt  # noqa: E402
    import graph as refgraph  # noqa: E402
    import networkx as nx  # noqa: E402

    mg.seed_all(0)
    dqn = mdt.MultiDismantler()
    dqn.LoadModel(os.path.join(vdir, mg.CKPT[variant]))
    net = dqn.MultiDismantler_net
    g_list, covered, actions, remove_edges, targets, meta = [], [], [], [], [], []
    cases = [p for p in sorted(glob.glob(os.path.join(HERE, "unit_n*_s*.npz"))) if int(np.load(p)["num_nodes"]) <= 60]
    for ci, p in enumerate(cases):
        z = np.load(p)
        n = int(z["num_nodes"])
        G1, G2 = nx.Graph(), nx.Graph()
        G1.add_nodes_from(range(n))
        G2.add_nodes_from(range(n))
        G1.add_edges_from((int(u), int(v)) for u, v in z["edges0"])
        G2.add_edges_from((int(u), int(v)) for u, v in z["edges1"])
        g = refgraph.Graph_test(G1, G2)
        assert [tuple(e) for e in g.edge_list[0]] == [tuple(int(x) for x in e) for e in z["edges0"]]
        acts = [int(a) for a in z["greedy_actions"]]
        for t in (0, min(2, len(acts) - 1)):
            env = dqn.test_env
            env.s0(g)
            for a in acts[:t]:
                env.stepWithoutReward(a)
            g_list.append(g)
            covered.append(list(env.action_list))
            actions.append(acts[t])
            remove_edges.append([set(env.remove_edge[0]), set(env.remove_edge[1])])
            targets.append([-0.05 - 0.01 * len(targets)])
            meta.append((os.path.basename(p)[:-4], t))
    B = len(g_list)
    idxes = np.arange(B, dtype=np.int32)
    dqn.optimizer.zero_grad()
    dqn.SetupTrain(idxes, g_list, covered, actions, np.array(targets), remove_edges)
    q_pred, cur = net.train_forward(node_input=dqn.inputs["node_input"], subgsum_param=dqn.inputs["subgsum_param"],
                                    n2nsum_param=dqn.inputs["n2nsum_param"], action_select=dqn.inputs["action_select"],
                                    aux_input=dqn.inputs["aux_input"], adj=dqn.inputs["adj"], v_adj=dqn.inputs["v_adj"])
    loss = dqn.calc_loss(q_pred, cur)
    loss.backward()
    out = {
        "cases": np.array([m[0] for m in meta]), "t": np.array([m[1] for m in meta], np.int32),
        "actions": np.array(actions, np.int32), "targets": np.array(targets, np.float64),
        "q_pred": q_pred.detach().numpy().astype(np.float32), "loss": np.float64(loss.item()),
        "emb0": cur[0].detach().numpy().astype(np.float32), "emb1": cur[1].detach().numpy().astype(np.float32),
        "alpha": np.float64(mdt.Alpha), "lr": np.float64(mdt.LEARNING_RATE),
    }
    seen = set()
    for name, prm in net.named_parameters():
        if id(prm) in seen:  # last_w is h2_weight (net_graphsage.py:69)
            continue
        seen.add(id(prm))
        out["grad/" + name] = prm.grad.detach().numpy().astype(np.float32).copy()
    dqn.optimizer.step()
    seen = set()
    for name, prm in net.named_parameters():
        if id(prm) in seen:
            continue
        seen.add(id(prm))
        out["after/" + name] = prm.detach().numpy().astype(np.float32).copy()
    np.savez_compressed(os.path.join(HERE, "train_unit.npz"), **out)
    print("batch", B, "loss %.10f" % loss.item(), "q_pred", out["q_pred"][:4, 0],
          "grad keys", sorted(k for k in out if k.startswith("grad/")))


if __name__ == "__main__":
    main()
