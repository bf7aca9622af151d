"""Config C4: one 3-layer synthetic multiplex with N = 1M nodes and 4M undirected edges per layer.
Times >= `--steps` consecutive greedy steps from the intact graph and reports, per step, the aggregation stage
(3 md_embed_iter launches = 9 layer-iterations), the LMCC / env stage and the whole step, with the SURVEY 8(d)
no-reuse byte model for the aggregation (H_l = 256 MB > L2)."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mdcommunity_b200 as mdb  # noqa: E402
from mdcommunity_b200 import _lib, gmm  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1_000_000)
ap.add_argument("--layers", type=int, default=3)
ap.add_argument("--edges", type=int, default=4_000_000)
ap.add_argument("--steps", type=int, default=100)
ap.add_argument("--profile-steps", type=int, default=0, help="bracket this many steps with cudaProfilerStart/Stop")
ap.add_argument("--dbg", type=int, default=0, help="md_debug_set_flags value (A/B switches of the kernels)")
args = ap.parse_args()

t0 = time.time()
layers = gmm.random_sparse_layers(args.n, args.layers, args.edges, seed=0)
g = mdb.Graph_test(*layers, num_nodes=args.n, cost="unit")
batch = mdb.GraphBatch([g], "cuda:0", "unit")
w = mdb.QWeights(dict(np.load(os.path.join(ROOT, "tests", "golden", "weights_unit.npz"))), batch.device)
ro = mdb.BatchRollout(batch, w, max_steps=args.steps + 16)
torch.cuda.synchronize()
print("setup %.1f s; max_rank %d, active %d, alive directed edges %s" % (
    time.time() - t0, g.max_rank, int(ro.n_act[0]), (2 * ro.alive_edges[:, 0]).tolist()), flush=True)
lib = ro.lib
lib.md_debug_set_flags(args.dbg)
gp, sp, wp, kp, st = ro._g(), C.byref(ro.s), ro.w.ptr(), C.byref(ro.k), ro._stream()
ev = lambda: torch.cuda.Event(enable_timing=True)
rows = []
for step in range(args.steps + 3):
    if args.profile_steps and step == 3:
        torch.cuda.profiler.start()
    e = [ev() for _ in range(5)]
    n, nnz = int(ro.n_act[0]), (2 * ro.alive_edges[:, 0]).cpu().numpy()
    e[0].record()
    _lib.check(lib.md_embed_init(gp, sp, wp, kp, st), "init")
    e[1].record()
    for it in range(3):
        _lib.check(lib.md_embed_iter(gp, sp, wp, kp, it, st), "iter")
    e[2].record()
    _lib.check(lib.md_graph_embed(gp, sp, wp, kp, st), "gembed")
    _lib.check(lib.md_qhead(gp, sp, wp, kp, st), "qhead")
    e[3].record()
    ro.step()
    e[4].record()
    torch.cuda.synchronize()
    if args.profile_steps and step == 2 + args.profile_steps:
        torch.cuda.profiler.stop()
    if step == 5:
        f6, f7 = int(ro.flags[6]), int(ro.flags[7])
        print("env_step_large: %d MCC rounds %.1f us, component sizes + alive degrees %d us" % (f7 >> 24, f6 / 1e3, f7 & 0xffffff), flush=True)
    if step >= 3:
        rows.append([e[i].elapsed_time(e[i + 1]) for i in range(4)] + [n, int(nnz.sum()), int(ro.rank[0])])
    if int(ro.done[0]):
        break
r = np.array(rows, dtype=np.float64)
med = np.median(r[:, :4], 0)
n, nnz = r[0, 4], r[0, 5]
L, D = args.layers, 64
spmm_bytes = 3 * (L * 4 * (n + 1) + 4 * nnz + 4 * D * nnz + L * 4 * D * n)  # 3 iterations x L layers, no-reuse model
floor_bytes = 3 * (L * 4 * (n + 1) + 4 * nnz + L * 8 * D * n)
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
out = {
    "config": "C4: N=%d, L=%d, %d undirected edges/layer, %d steps from the intact graph" % (args.n, L, args.edges, len(r)),
    "median_ms": {"embed_init": med[0], "aggregation_3iters": med[1], "graph_embed+qhead": med[2], "env_step_lmcc": med[3],
                  "step_total": float(np.median(r[:, :4].sum(1)))},
    "removal_steps_per_s": 1e3 / float(np.median(r[:, :4].sum(1))),
    "aggregation": {"no_reuse_bytes": spmm_bytes, "floor_bytes": floor_bytes,
                    "achieved_gbs_no_reuse": spmm_bytes / (med[1] * 1e-3) / 1e9,
                    "achieved_gbs_floor": floor_bytes / (med[1] * 1e-3) / 1e9, "peak_gbs": peak,
                    "frac_no_reuse": spmm_bytes / (med[1] * 1e-3) / 1e9 / peak},
    "lmcc_last_rank": int(r[-1, 6]), "active_nodes": int(n),
}
print(json.dumps(out))
