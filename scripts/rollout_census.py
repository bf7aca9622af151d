"""Census of a C2 rollout: at every lock-step iteration the number of live graphs, the distribution of active nodes per
graph, the per-graph cycle counts / prune rounds of the env step and (every `--every` iterations) the phase clocks of
md_embed_fused.  Planning data for the scheduler / kernel work; prints one line per iteration."""
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
ap.add_argument("--every", type=int, default=8)
ap.add_argument("--pack", type=int, default=0, help="md_debug_set_fused_pack: 0 = default budget, > 0 fixed, < 0 no packing")
ap.add_argument("--legacy", action="store_true", help="md_embed_fused as it was before small units were packed (debug bit 19)")
args = ap.parse_args()
_lib.lib().md_debug_set_fused_pack(args.pack)
if args.legacy:
    _lib.lib().md_debug_set_flags(1 << 19)

graphs = [gmm.gmm_graph(args.n, i, cost=args.cost) for i in range(args.graphs)]
batch = mdb.GraphBatch(graphs, "cuda:0", args.cost)
w = mdb.QWeights(dict(np.load(os.path.join(ROOT, "tests", "golden", "weights_%s.npz" % args.cost))), batch.device)
ro = mdb.BatchRollout(batch, w)
ro.run()
ro.reset()
lib = ro.lib
torch.cuda.synchronize()
it = 0
ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
while ro.num_done() < batch.B and it < 200:
    nact = ro.n_act.cpu().numpy()
    done = ro.done.cpu().numpy()
    live = nact[done == 0]
    g, s, wp, k, st = ro._g(), C.byref(ro.s), ro.w.ptr(), C.byref(ro.k), ro._stream()
    ev[0].record()
    _lib.check(lib.md_embed_fused(g, s, wp, k, st), "fused")
    ev[1].record()
    _lib.check(lib.md_graph_embed(g, s, wp, k, st), "gembed")
    ev[2].record()
    _lib.check(lib.md_qhead(g, s, wp, k, st), "qhead")
    ev[3].record()
    ro.step()
    ev[4].record()
    torch.cuda.synchronize()
    ms = [1e3 * ev[i].elapsed_time(ev[i + 1]) for i in range(4)]
    cb = np.zeros((1024, 4), np.uint32)
    lib.md_debug_read_env_ctas(cb.ctypes.data_as(C.c_void_p))
    cb = cb[:args.graphs][done == 0]
    cyc = cb[:, 0].astype(np.int64)
    line = "it %3d live %3d n_act[min/med/max] %4d %4d %4d  n<=512: %3d n<=256: %3d n<=128: %3d | us fused %4.0f gemb %3.0f qhead %4.0f env %4.0f | env cyc med %6d p90 %6d max %6d rounds med %d max %d" % (
        it, len(live), live.min(), np.median(live), live.max(), (live <= 512).sum(), (live <= 256).sum(), (live <= 128).sum(),
        ms[0], ms[1], ms[2], ms[3], np.median(cyc), np.percentile(cyc, 90), cyc.max(), np.median(cb[:, 1]), cb[:, 1].max())
    print(line)
    if it % args.every == 0:
        buf = np.zeros(32, np.int64)
        lib.md_debug_read_env(buf.ctypes.data_as(C.c_void_p), 32)
        print("    env CTA0 phases [labels, stage, prune, sizes, aliveCSR]:", buf[:5].tolist(), "rounds", int(buf[16]), "hook/flatten", buf[17:19].tolist())
        fb = np.zeros(16, np.int64)
        lib.md_debug_read_fused(fb.ctypes.data_as(C.c_void_p), 16)
        print("    fused CTA0 gather warp [setup, fill, reload, gather, stage wait, pass-end, epi math, idle]", fb[:8].tolist(), "[epi stores, epi sync]", fb[12:14].tolist(),
              "MMA [Wwait, full wait, issue, completion]", fb[8:12].tolist(), "units", int(fb[15]))
        gb = np.zeros(8, np.uint32)
        lib.md_debug_read_fused_gather(gb.ctypes.data_as(C.c_void_p))
        print("    gather sub [row setup, nbr loop, own rows, stores, long rows, fence]", gb[:6].tolist())
        fc = np.zeros((160, 2), np.uint32)
        lib.md_debug_read_fused_ctas(fc.ctypes.data_as(C.c_void_p))
        fc = fc[fc[:, 1] > 0]
        if len(fc):
            print("    fused per-CTA cycles min %d mean %d max %d; units/CTA min %d max %d; cycles/unit %d" % (
                fc[:, 0].min(), fc[:, 0].mean(), fc[:, 0].max(), fc[:, 1].min(), fc[:, 1].max(), fc[:, 0].sum() / fc[:, 1].sum()))
    it += 1
res = ro.results()
print("steps per graph: min %d mean %.1f max %d" % (res["n_steps"].min(), res["n_steps"].mean(), res["n_steps"].max()))
print("hist of steps:", np.histogram(res["n_steps"], bins=[0, 10, 20, 30, 40, 50, 60, 70, 80, 200])[0].tolist())
