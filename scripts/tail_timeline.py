"""Timeline of a C2 rollout with the tail hand-over: per chunk the GPU time of the lock-step iterations (events on the
main stream) and the total, for several hand-over policies.  Planning data for the tail scheduling."""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mdcommunity_b200 as mdb  # noqa: E402
from mdcommunity_b200 import _lib, gmm  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cost", default="degree")
ap.add_argument("--n", type=int, default=1000)
ap.add_argument("--graphs", type=int, default=256)
ap.add_argument("--reps", type=int, default=5)
args = ap.parse_args()

graphs = [gmm.gmm_graph(args.n, i, cost=args.cost) for i in range(args.graphs)]
batch = mdb.GraphBatch(graphs, "cuda:0", args.cost)
w = mdb.QWeights(dict(np.load(os.path.join(ROOT, "tests", "golden", "weights_%s.npz" % args.cost))), batch.device)


def timed_run(ro, chunk=8, defer=0.0, verbose=False):
    ro.reset()
    torch.cuda.synchronize()
    B = ro.batch.B
    ev = [torch.cuda.Event(enable_timing=True)]
    ev[0].record()
    s = ro._run_state()
    lib = ro.lib
    st = ro._stream()
    if ro.tail:
        _lib.check(lib.md_tail_handover(ro._g(), C.byref(ro.s_tail), st), "handover")
        if defer <= 0:
            _lib.check(lib.md_rollout_tail(ro._g(), C.byref(ro.s_tail), ro.w.ptr(), st), "tail")
    left_hist = []
    it = 0
    left = 0
    while True:
        _lib.check(lib.md_rollout_steps(ro._g(), C.byref(s), ro.w.ptr(), C.byref(ro.k), chunk, st), "steps")
        if ro.tail and left >= defer * B:
            _lib.check(lib.md_rollout_tail(ro._g(), C.byref(ro.s_tail), ro.w.ptr(), st), "tail")
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        ev.append(e)
        it += chunk
        left = ro.num_left()
        left_hist.append(left)
        if left >= B or it > ro.max_steps + chunk:
            break
    e_ls = torch.cuda.Event(enable_timing=True)
    e_ls.record()
    if ro.tail:
        _lib.check(lib.md_rollout_tail(ro._g(), C.byref(ro.s_tail), ro.w.ptr(), st), "tail")
        _lib.check(lib.md_tail_join(st), "join")
    e_end = torch.cuda.Event(enable_timing=True)
    e_end.record()
    torch.cuda.synchronize()
    total = ev[0].elapsed_time(e_end)
    if verbose:
        ts = [ev[0].elapsed_time(x) for x in ev[1:]]
        print("   chunk end times (ms):", [round(t, 2) for t in ts])
        print("   left after chunk:", left_hist, " lock-step end %.2f ms, total %.2f ms" % (ev[0].elapsed_time(e_ls), total))
    assert ro.num_done() == B
    return total


for name, kw, defer in (("lock-step only", dict(tail=False), 0.0), ("tail, launch per chunk", dict(tail=True), 0.0),
                        ("tail, launches deferred until 50% left", dict(tail=True), 0.5),
                        ("tail, launches deferred until 90% left", dict(tail=True), 0.9),
                        ("tail, only at the end", dict(tail=True), 2.0)):
    ro = mdb.BatchRollout(batch, w, **kw)
    ro.run()
    ts = [timed_run(ro, defer=defer, verbose=(r == 0)) for r in range(args.reps)]
    print("%-45s median %.2f ms  (min %.2f)  mean removals %.1f" % (name, np.median(ts), min(ts), float(ro.n_steps.float().mean())))

# pure tail cost: hand over everything that qualifies at s0 of small graphs
small = [gmm.gmm_graph(100, 500 + i, cost=args.cost) for i in range(args.graphs)]
bs = mdb.GraphBatch(small, "cuda:0", args.cost)
for kw in (dict(tail=False), dict(tail=True)):
    ro = mdb.BatchRollout(bs, w, **kw)
    ro.run()
    ts = []
    for r in range(args.reps):
        ro.reset()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ro.run()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    steps = float(ro.n_steps.float().sum())
    print("256 graphs of N=100, %s: %.2f ms per rollout, %.0f graph-steps -> %.1f us per graph-step per SM-equivalent" % (
        kw, np.median(ts), steps, 1e3 * np.median(ts) * min(148, args.graphs) / steps))
