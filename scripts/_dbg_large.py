import os, sys, numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import mdcommunity_b200 as mdb
from mdcommunity_b200 import _lib, gmm
lib = _lib.lib()
cost = "unit"
graphs = [gmm.gmm_graph(n, 7 + i, cost=cost) for i, n in enumerate((60, 120, 200, 300))]
w = None
res = {}
for mode in (0, 1):
    lib.md_debug_force_large_env(mode)
    batch = mdb.GraphBatch(graphs, "cuda:0", cost)
    w = mdb.QWeights(dict(np.load("/root/repo/tests/golden/weights_%s.npz" % cost)), batch.device)
    ro = mdb.BatchRollout(batch, w)
    r = ro.run()
    res[mode] = r
    print("mode", mode, "large", ro.large, "n_steps", r["n_steps"], "score", r["score"])
lib.md_debug_force_large_env(0)
for i in range(len(graphs)):
    a, b = res[0], res[1]
    n = min(a["n_steps"][i], b["n_steps"][i])
    d = np.nonzero(a["actions"][i, :n] != b["actions"][i, :n])[0]
    print(i, "first action diff", d[:1], "ranks equal", np.array_equal(a["ranks"][i, :n], b["ranks"][i, :n]))
    if len(d):
        t = d[0]; print("  small", a["actions"][i, t-1:t+2], a["ranks"][i, t-1:t+2], " large", b["actions"][i, t-1:t+2], b["ranks"][i, t-1:t+2])
