"""Profiling target for md_rollout_tail: 256 graphs of N = 100 go to the tail at s0, so one rollout = one tail launch that
does all the work.  Use: ncu --set full -k regex:rollout_tail --launch-skip 2 --launch-count 1 python scripts/profile_tail.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mdcommunity_b200 as mdb  # noqa: E402
from mdcommunity_b200 import gmm  # noqa: E402

cost = sys.argv[1] if len(sys.argv) > 1 else "degree"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100
graphs = [gmm.gmm_graph(n, 500 + i, cost=cost) for i in range(256)]
batch = mdb.GraphBatch(graphs, "cuda:0", cost)
w = mdb.QWeights(dict(np.load(os.path.join(ROOT, "tests", "golden", "weights_%s.npz" % cost))), batch.device)
ro = mdb.BatchRollout(batch, w)
for rep in range(3):
    ro.reset()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    ro.run()
    b.record()
    torch.cuda.synchronize()
    print("rollout %.2f ms, %d graph-steps" % (a.elapsed_time(b), int(ro.n_steps.sum())))

import ctypes as C
buf = np.zeros((1024, 10), np.uint64)
ro.lib.md_debug_read_tail(buf.ctypes.data_as(C.c_void_p))
b = buf[:256].astype(np.int64)
if b[:, 0].max() > 0:
    st = np.maximum(b[:, 1], 1)
    print("tail per graph: cycles min %d median %d max %d; steps min %d mean %.1f max %d; cycles/step median %d p90 %d" % (
        b[:, 0].min(), np.median(b[:, 0]), b[:, 0].max(), b[:, 1].min(), b[:, 1].mean(), b[:, 1].max(),
        np.median(b[:, 0] / st), np.percentile(b[:, 0] / st, 90)))
    tot = b[:, 4:].sum(0)
    print("phase share [setup, embed iters, graph rows, qhead+argmax, LMCC rounds, env rest]:", (tot / tot.sum()).round(3).tolist(),
          " per step cycles:", (tot / b[:, 1].sum()).round(0).tolist())
    sm, cnt = np.unique(b[:, 3], return_counts=True)
    load = np.array([b[b[:, 3] == x, 0].sum() for x in sm])
    print("SMs used %d, graphs per SM max %d, busiest SM %d cycles, mean %d" % (len(sm), cnt.max(), load.max(), load.mean()))
