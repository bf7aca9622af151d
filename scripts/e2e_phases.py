"""Host-side phase times of one end-to-end C2 batch: GraphBatch (H2D), BatchRollout (state allocation + s0), run, results (D2H)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mdcommunity_b200 as mdb  # noqa: E402
from mdcommunity_b200 import gmm  # noqa: E402

cost, n, B = "degree", 1000, 256
graphs = [gmm.gmm_graph(n, i, cost=cost) for i in range(B)]
dev = torch.device("cuda:0")
w = mdb.QWeights(dict(np.load(os.path.join(ROOT, "tests", "golden", "weights_%s.npz" % cost))), dev)
host = mdb.GraphBatch.prepare_host(graphs, cost)
for rep in range(8):
    torch.cuda.synchronize()
    t = [time.perf_counter()]
    b = mdb.GraphBatch(graphs, dev, cost, host=host)
    torch.cuda.synchronize(); t.append(time.perf_counter())
    r = mdb.BatchRollout(b, w)
    torch.cuda.synchronize(); t.append(time.perf_counter())
    r.run()
    torch.cuda.synchronize(); t.append(time.perf_counter())
    out = r.results()
    torch.cuda.synchronize(); t.append(time.perf_counter())
    d = np.diff(t) * 1e3
    print("rep %d: GraphBatch %.2f ms, BatchRollout ctor+s0 %.2f ms, run %.2f ms, results %.2f ms, total %.2f ms -> %.0f graphs/s" % (
        rep, d[0], d[1], d[2], d[3], d.sum(), B / d.sum() * 1e3))
