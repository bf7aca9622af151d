"""Experiment: C2 rollouts of several independent sub-batches interleaved on separate CUDA streams (tails and
latency-bound kernels of one sub-batch overlap the other's), against the single-stream lock-step rollout."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mdcommunity_b200 as mdb  # noqa: E402
from mdcommunity_b200 import _lib, gmm  # noqa: E402

cost, n, B = "degree", 1000, 256
graphs = [gmm.gmm_graph(n, i, cost=cost) for i in range(B)]
dev = torch.device("cuda:0")
w = mdb.QWeights(dict(np.load(os.path.join(ROOT, "tests", "golden", "weights_%s.npz" % cost))), dev)


def make(parts):
    ros, streams = [], []
    per = B // parts
    for p in range(parts):
        batch = mdb.GraphBatch(graphs[p * per:(p + 1) * per], dev, cost)
        ros.append(mdb.BatchRollout(batch, w))
        streams.append(torch.cuda.Stream(dev))
    return ros, streams


def run_multi(ros, streams, chunk=8):
    pins = [torch.zeros(1, dtype=torch.int32).pin_memory() for _ in ros]
    evs = [torch.cuda.Event() for _ in ros]
    for ro, st in zip(ros, streams):
        with torch.cuda.stream(st):
            ro.reset()
    live = list(range(len(ros)))
    its = 0
    while live:
        for i in live:
            ro = ros[i]
            with torch.cuda.stream(streams[i]):
                _lib.check(ro.lib.md_rollout_steps(ro._g(), C.byref(ro.s), ro.w.ptr(), C.byref(ro.k), chunk, ro._stream()), "steps")
                pins[i].copy_(ro.n_done, non_blocking=True)
                evs[i].record(streams[i])
        nxt = []
        for i in live:
            evs[i].synchronize()
            if int(pins[i][0]) < ros[i].batch.B:
                nxt.append(i)
        live = nxt
        its += chunk
        if its > 2000:
            break
    torch.cuda.synchronize()


for parts in (1, 2, 4):
    ros, streams = make(parts)
    for _ in range(3):
        run_multi(ros, streams)
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        run_multi(ros, streams)
    dt = (time.perf_counter() - t0) / reps
    sc = np.concatenate([ro.score.cpu().numpy() for ro in ros])
    print("%d sub-batch(es) of %d graphs on %d stream(s): %.2f ms per 256 graphs = %.0f graphs/s  (mean AUDC %.6f)" % (
        parts, B // parts, parts, 1e3 * dt, B / dt, sc.mean()))
