"""One complete C2 rollout bracketed by cudaProfilerStart/Stop (ncu --profile-from-start off launch list)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mdcommunity_b200 as mdb  # noqa: E402
from mdcommunity_b200 import gmm  # noqa: E402

cost, n, B = "degree", 1000, int(sys.argv[1]) if len(sys.argv) > 1 else 256
graphs = [gmm.gmm_graph(n, i, cost=cost) for i in range(B)]
batch = mdb.GraphBatch(graphs, "cuda:0", cost)
w = mdb.QWeights(dict(np.load(os.path.join(ROOT, "tests", "golden", "weights_%s.npz" % cost))), batch.device)
ro = mdb.BatchRollout(batch, w)
ro.run()
ro.reset()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.profiler.start()
e0.record()
res = ro.run()
e1.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("rollout of %d graphs: %.2f ms, max steps %d, mean %.1f" % (B, e0.elapsed_time(e1), res["n_steps"].max(), res["n_steps"].mean()))
