"""How far the device's Q is from the reference's own f32 Q on the golden greedy trajectories (teacher-forced), per
(cost variant, embedding path): largest absolute error, largest error relative to the tolerance 1e-4 |q| + 1e-6
(1.0 = at the bar), over every step of every golden graph.  The parity test asserts <= 1.0; this prints the margin."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mdcommunity_b200 as mdb  # noqa: E402
from conftest import GOLDEN, golden_cases  # noqa: E402

Q_RTOL, Q_ATOL = 1e-4, 1e-6
for variant in ("unit", "degree"):
    zs, graphs = [], []
    for p in golden_cases(variant, None):
        z = np.load(p)
        zs.append(z)
        graphs.append(mdb.Graph_test(z["edges0"], z["edges1"], num_nodes=int(z["num_nodes"]), cost=variant))
    batch = mdb.GraphBatch(graphs, cost=variant)
    w = mdb.QWeights(dict(np.load(os.path.join(GOLDEN, "weights_%s.npz" % variant))), batch.device)
    T = max(len(z["greedy_actions"]) for z in zs)
    acts = np.full((len(zs), T + 1), -1, np.int32)
    for i, z in enumerate(zs):
        acts[i, :len(z["greedy_actions"])] = z["greedy_actions"]
    acts = torch.from_numpy(acts).to(batch.device)
    for fused in (True, False):
        ro = mdb.BatchRollout(batch, w, fused=fused)
        worst_abs = worst_ratio = 0.0
        for t in range(T):
            ro.forward()
            for i, z in enumerate(zs):
                if t >= len(z["greedy_actions"]):
                    continue
                off = z["greedy_off"]
                ids, q = ro.q_values(i)
                qg = z["greedy_q"][off[t]:off[t + 1]]
                e = np.abs(q - qg)
                worst_abs = max(worst_abs, float(e.max()))
                worst_ratio = max(worst_ratio, float((e / (Q_RTOL * np.abs(qg) + Q_ATOL)).max()))
            ro.step(acts)
        print("%-6s %-8s worst |dq| %.3e   worst |dq| / (1e-4 |q| + 1e-6) = %.3f" % (variant, "fused" if fused else "unfused", worst_abs, worst_ratio))
