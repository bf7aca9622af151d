#!/usr/bin/env python
"""Golden vectors for the heuristic baselines: runs the UNMODIFIED reference scripts
baseline/HDA/hda_2max.py and baseline/CI/ci_max.py (critical_number) on the committed golden graphs
(tests/golden/unit_n*_s*.npz) and records the removal sequence, the normalised LMCC list and the AUDC score.

Only this script imports the reference; it cannot run on the GPU box.  Output: tests/golden/baselines_unit.npz.
Shims: a matplotlib stub (the scripts import pyplot / ticker at module import); nothing else is touched.

    python tests/golden/make_golden_baselines.py
"""
import glob
import importlib.util
import os
import sys
import types

import networkx as nx
import numpy as np

REF = "/root/reference/code/MultiDismantler_unit_cost/baseline"
OUT = os.path.dirname(os.path.abspath(__file__))


def install_shims():
    mpl = types.ModuleType("matplotlib")
    mpl.use = lambda *a, **k: None
    plt = types.ModuleType("matplotlib.pyplot")
    tick = types.ModuleType("matplotlib.ticker")
    tick.MultipleLocator = tick.FuncFormatter = object
    mpl.pyplot, mpl.ticker = plt, tick
    sys.modules.update({"matplotlib": mpl, "matplotlib.pyplot": plt, "matplotlib.ticker": tick})


def load(path, name):
    cwd = os.getcwd()
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)  # the scripts os.chdir(sys.path[0]) at import
    os.chdir(cwd)
    return mod


def build(n, edges):
    g = nx.Graph()
    g.add_nodes_from(range(n))
    g.add_edges_from((int(u), int(v)) for u, v in edges)
    return g


def main():
    install_shims()
    hda = load(os.path.join(REF, "HDA", "hda_2max.py"), "ref_hda_2max")
    ci = load(os.path.join(REF, "CI", "ci_max.py"), "ref_ci_max")
    out = {}
    for p in sorted(glob.glob(os.path.join(OUT, "unit_n*_s*.npz"))):
        z = np.load(p)
        n = int(z["num_nodes"])
        if n > 200:
            continue
        case = os.path.basename(p)[:-4]
        G1, G2 = build(n, z["edges0"]), build(n, z["edges1"])
        # as in the scripts' __main__ (hda_2max.py:150-151): M_ori = MCC(G1, G2) prunes G1 / G2 IN PLACE, the rollouts
        # then start from copies of the pruned layers
        m_ori = hda.MCC(G1, G2)
        assert m_ori == int(z["max_rank"])
        mcc, sol, score = hda.critical_number(G1.copy(), G2.copy(), n, m_ori)
        out[case + "/hda_sol"] = np.array(sol, np.int32)
        out[case + "/hda_mcc"] = np.array(mcc, np.float64)
        out[case + "/hda_score"] = np.float64(score)
        mcc, sol = ci.critical_number(G1.copy(), G2.copy(), n, m_ori)
        out[case + "/ci_sol"] = np.array(sol, np.int32)
        out[case + "/ci_mcc"] = np.array(mcc, np.float64)
        print(case, "HDA steps", len(out[case + "/hda_sol"]), "score %.10f" % score, "CI steps", len(sol))
    np.savez_compressed(os.path.join(OUT, "baselines_unit.npz"), **out)


if __name__ == "__main__":
    main()
