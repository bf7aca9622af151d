"""Experiment: C2 rollouts of several independent sub-batches interleaved on separate CUDA streams (MultiStreamRollout:
tails and latency-bound kernels of one sub-batch overlap the other's), against the single-stream lock-step rollout."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mdcommunity_b200 as mdb  # noqa: E402
from mdcommunity_b200 import gmm  # noqa: E402
from mdcommunity_b200.rollout import MultiStreamRollout  # noqa: E402

cost, n, B = "degree", 1000, 256
graphs = [gmm.gmm_graph(n, i, cost=cost) for i in range(B)]
dev = torch.device("cuda:0")
w = mdb.QWeights(dict(np.load(os.path.join(ROOT, "tests", "golden", "weights_%s.npz" % cost))), dev)
base = None
for parts in (1, 2, 3, 4):
    mr = MultiStreamRollout(graphs, w, dev, cost, parts=parts)
    for _ in range(3):
        r = mr.run()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 8
    for _ in range(reps):
        r = mr.run()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    if base is None:
        base = r
    same = np.array_equal(base["score"], r["score"]) and np.array_equal(base["actions"][:, :r["actions"].shape[1]], r["actions"][:, :base["actions"].shape[1]])
    print("%d sub-batch(es) on %d stream(s): %.2f ms per 256 graphs = %.0f graphs/s (wall clock incl. results D2H); identical results: %s" % (
        parts, parts, 1e3 * dt, B / dt, same))
