"""Heuristic baselines (HDA / CI) on the device-resident environment.

Mirror of the reference's baseline scripts (/root/reference/code/MultiDismantler_unit_cost/baseline/HDA/hda_2max.py,
hda_add.py, CI/ci_max.py, ci_add.py): `critical_number(G1, G2, N, ND_ori) -> (ND_mcc, solutions, score)`.
Selection, node removal, LMCC recompute and AUDC all run in the CUDA kernels (md_heuristic_select + md_env_step);
a whole list of graphs is dismantled in one lock-step batch.
"""
import numpy as np

from .graph import GraphBatch, Graph_test
from .rollout import BatchRollout


def critical_number_batch(graphs, kind="hda", combine="max", device=None):
    """graphs: list of Graph_test.  Returns a list of (ND_mcc, solutions, score) exactly shaped like the scripts' output:
    ND_mcc = [1, LMCC_1 / LMCC_0, ...], solutions = removed node ids, score = sum_t LMCC_t / (LMCC_0 * N)."""
    batch = GraphBatch(graphs, device, "unit") if device is not None else GraphBatch(graphs, cost="unit")
    ro = BatchRollout(batch, None, fused=False)
    res = ro.run_heuristic(kind, combine)
    out = []
    for i, g in enumerate(graphs):
        n = int(res["n_steps"][i])
        sol = [int(a) for a in res["actions"][i, :n]]
        nd = [1] + [float(r) / g.max_rank for r in res["ranks"][i, :n]]
        out.append((nd, sol, float(res["score"][i])))
    return out


def critical_number(G1, G2, N, ND_ori=None, kind="hda", combine="max", device=None):
    """Single-graph form with the scripts' argument order; G1 / G2 are edge arrays [E, 2] (or anything np.asarray
    turns into one).  ND_ori is recomputed on the device (it is the LMCC of the intact pair) and checked if given."""
    g = Graph_test(np.asarray(G1).reshape(-1, 2), np.asarray(G2).reshape(-1, 2), num_nodes=N, cost="unit")
    (nd, sol, score), = critical_number_batch([g], kind, combine, device)
    if ND_ori is not None and int(ND_ori) != g.max_rank:
        raise ValueError("ND_ori %d does not match the LMCC of the intact graph %d" % (ND_ori, g.max_rank))
    return nd, sol, score
