"""Short profiling target: a warm-up rollout, then `--steps` lock-step iterations from the intact state bracketed by
cudaProfilerStart/Stop (use ncu --profile-from-start off).  Prints CUDA-event time of the bracketed region."""
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
ap.add_argument("--steps", type=int, default=4)
ap.add_argument("--skip", type=int, default=0, help="un-profiled steps after the reset")
ap.add_argument("--dbg", type=int, default=0, help="md_debug_set_flags value for the measured region")
ap.add_argument("--iters", type=int, default=0, help="time N md_embed_iter launches alone instead of full steps")
ap.add_argument("--fused-cmp", type=int, default=0, help="time N embedding stages: md_embed_fused vs md_embed_init + 3 md_embed_iter")
ap.add_argument("--unfused", action="store_true", help="use the unfused embedding kernels in the rollout")
args = ap.parse_args()

graphs = [gmm.gmm_graph(args.n, i, cost=args.cost) for i in range(args.graphs)]
batch = mdb.GraphBatch(graphs, "cuda:0", args.cost)
w = mdb.QWeights(dict(np.load(os.path.join(ROOT, "tests", "golden", "weights_%s.npz" % args.cost))), batch.device)
ro = mdb.BatchRollout(batch, w, fused=False if args.unfused else None)
ro.run()
ro.reset()
lib = ro.lib
if args.skip:
    _lib.check(lib.md_rollout_steps(ro._g(), C.byref(ro.s), ro.w.ptr(), C.byref(ro.k), args.skip, ro._stream()), "steps")
torch.cuda.synchronize()
print("active nodes", int(ro.n_act.sum()), "alive directed edges", (2 * ro.alive_edges.sum(1)).tolist())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
if args.fused_cmp:
    for name in ("fused", "unfused"):
        for rep in range(3):
            torch.cuda.synchronize()
            if rep == 2 and name == "fused":
                torch.cuda.profiler.start()
            e0.record()
            for i in range(args.fused_cmp):
                if name == "fused":
                    _lib.check(lib.md_embed_fused(ro._g(), C.byref(ro.s), ro.w.ptr(), C.byref(ro.k), ro._stream()), "fused")
                else:
                    _lib.check(lib.md_embed_init(ro._g(), C.byref(ro.s), ro.w.ptr(), C.byref(ro.k), ro._stream()), "init")
                    for it in range(3):
                        _lib.check(lib.md_embed_iter(ro._g(), C.byref(ro.s), ro.w.ptr(), C.byref(ro.k), it, ro._stream()), "iter")
            e1.record()
            torch.cuda.synchronize()
            if rep == 2 and name == "fused":
                torch.cuda.profiler.stop()
        print("%s: %.1f us per embedding stage (first embedding + 3 iterations)" % (name, 1e3 * e0.elapsed_time(e1) / args.fused_cmp))
        if name == "fused":
            buf = np.zeros(16, np.int64)
            lib.md_debug_read_fused(buf.ctypes.data_as(C.c_void_p), 16)
            print("fused CTA0 cycles: gather warp [setup, fill, reload, gather, stage wait, pass-end wait, epi TMEM+math, idle] =", buf[:8].tolist(),
                  "[epi stores, epi fence+sync] =", buf[12:14].tolist(),
                  "\n   MMA warp [W+resident wait, issue, full-tile wait, completion wait] =", buf[8:12].tolist(), "; units", int(buf[15]))
            gb = np.zeros(8, np.uint32)
            lib.md_debug_read_fused_gather(gb.ctypes.data_as(C.c_void_p))
            print("   gather sub-phases [row setup, neighbour loop, own rows, stores, long rows, fence+arrive] =", gb[:6].tolist())
            cb = np.zeros((160, 2), np.uint32)
            lib.md_debug_read_fused_ctas(cb.ctypes.data_as(C.c_void_p))
            cb = cb[cb[:, 1] > 0]
            print("   per-CTA loop cycles: min %d mean %d max %d; units/CTA min %d max %d; cycles/unit mean %d" % (
                cb[:, 0].min(), cb[:, 0].mean(), cb[:, 0].max(), cb[:, 1].min(), cb[:, 1].max(), cb[:, 0].sum() / cb[:, 1].sum()))
    sys.exit(0)
if args.iters:
    _lib.check(lib.md_embed_init(ro._g(), C.byref(ro.s), ro.w.ptr(), C.byref(ro.k), ro._stream()), "init")
    for dbg in ((args.dbg,) if args.dbg else (0, 1, 2, 3)):
        lib.md_debug_set_flags(dbg)
        for rep in range(2):
            torch.cuda.synchronize()
            if rep == 1:
                torch.cuda.profiler.start()
            e0.record()
            for i in range(args.iters):
                _lib.check(lib.md_embed_iter(ro._g(), C.byref(ro.s), ro.w.ptr(), C.byref(ro.k), i % 3, ro._stream()), "iter")
            e1.record()
            torch.cuda.synchronize()
            if rep == 1:
                torch.cuda.profiler.stop()
        print("dbg=%d: %.1f us per md_embed_iter launch" % (dbg, 1e3 * e0.elapsed_time(e1) / args.iters))
    lib.md_debug_set_flags(0)
    sys.exit(0)
lib.md_debug_set_flags(args.dbg)
torch.cuda.profiler.start()
e0.record()
_lib.check(lib.md_rollout_steps(ro._g(), C.byref(ro.s), ro.w.ptr(), C.byref(ro.k), args.steps, ro._stream()), "steps")
e1.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("%d steps: %.3f ms (%.3f ms/step)" % (args.steps, e0.elapsed_time(e1), e0.elapsed_time(e1) / args.steps))
buf = np.zeros(32, np.int64)
lib.md_debug_read_env(buf.ctypes.data_as(C.c_void_p), 32)
print("env CTA0 phase cycles [labels, stage CSR, prune rounds, comp sizes, alive CSR+deg]:", buf[:5].tolist(), "rounds", int(buf[16]), "hook / flatten cycles", buf[17:19].tolist())
cb = np.zeros((1024, 4), np.uint32)
lib.md_debug_read_env_ctas(cb.ctypes.data_as(C.c_void_p))
cb = cb[:args.graphs]
cb = cb[cb[:, 0] > 0]
if len(cb):
    order = np.argsort(-cb[:, 0].astype(np.int64))[:5]
    print("env per-graph cycles: min %d median %d max %d; rounds min %d max %d; slowest (cycles, rounds, entries, sm):" % (
        cb[:, 0].min(), np.median(cb[:, 0]), cb[:, 0].max(), cb[:, 1].min(), cb[:, 1].max()), cb[order].tolist())
    sm, cnt = np.unique(cb[:, 3], return_counts=True)
    print("graphs per SM: max %d, SMs used %d" % (cnt.max(), len(sm)))
