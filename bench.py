#!/usr/bin/env python
"""Benchmark of the greedy-dismantling rollout (BASELINE.json metric: dismantled graphs/sec, N=1000, 2-layer).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c1|c3]

A "step" is one complete lock-step greedy rollout (s0 -> every graph terminal) of one batch of synthetic graphs.
Workload at N=1 (BASELINE.json configs[1], "c2"): degree-cost rollout, 256 two-layer GMM graphs of N=1000 per GPU.
With --gpus N (torchrun, one rank per GPU) every rank owns its own 256 graphs (weak scaling, no data-path
collective); the per-graph records are all-gathered over NCCL at the end of every step inside the timed region.

One JSON line is printed by rank 0 (see the contract in the task statement); extra keys: roofline, cpu_baseline.
`--impl reference` times the CPU restatement of the reference path (oracle/, kind "port": the reference is pure
Python and /root/reference does not exist on the GPU box) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (cost, N, graphs per GPU, description)
    "c2": ("degree", 1000, 256, "MultiDismantler_degree_cost rollout: 2-layer synthetic multiplex (GMM kbar=6) N=1000, batch of 256 graphs per GPU"),
    "c3": ("unit", 1000, 128, "unit-cost rollout: 2-layer synthetic GMM graphs N=1000 (kbar~U(2,10)), 128 graphs per GPU"),
    "c1": ("unit", 100, 1, "MultiDismantler_unit_cost testSynthetic: 2-layer synthetic multiplex N=100, 1 graph"),
}


def make_graphs(cost, n, count, seed0):
    from mdcommunity_b200 import gmm
    return [gmm.gmm_graph(n, seed0 + i, cost=cost) for i in range(count)]


def load_state_dict(cost):
    z = np.load(os.path.join(ROOT, "tests", "golden", "weights_%s.npz" % cost))
    return {k: z[k] for k in z.files}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, val in zip(names, f[3:7]):
                if val == "Active":
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores
# --------------------------------------------------------------------------------------------------------------
def oracle_rollouts(cost, graphs):
    from oracle import md_oracle as orc
    w = orc.load_weights(os.path.join(ROOT, "tests", "golden", "weights_%s.npz" % cost))
    steps = 0
    t0 = time.perf_counter()
    for g in graphs:
        og = orc.MultiplexGraph(g.num_nodes, g.edge_list, weights=g.weights if cost == "degree" else None)
        r = orc.rollout(w, og, cost)
        steps += len(r["sol"])
    return time.perf_counter() - t0, steps


def run_reference(args, cost, n, desc):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(min(16, cores))  # the reference sets 16 (MultiDismantler_torch.py:108)
    per_step = 2 if n >= 1000 else 8
    graphs = make_graphs(cost, n, per_step * (args.steps + args.warmup), 10_000)
    times = []
    for k in range(args.steps + args.warmup):
        dt, _ = oracle_rollouts(cost, graphs[k * per_step:(k + 1) * per_step])
        if k >= args.warmup:
            times.append(dt)
    total = sum(times)
    v = per_step * len(times) / total
    sample = "%d graphs per step (same generator/weights as the GPU arm), 1 process, torch threads=%d" % (
        per_step, torch.get_num_threads())
    out = {"impl": "reference", "metric": "dismantled graphs/sec", "value": v, "unit": "graphs/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": desc, "sample": sample},
           "cpu_baseline": {"value": v, "unit": "graphs/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sample,
                            "host_cores": cores},
           "e2e": {"value": v, "unit": "graphs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


# --------------------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------------------
def roofline_pass(mdb, batch, w):
    """Instrumented replay of one rollout with CUDA events around the dominant kernel of every step.

    Fused path (stride <= 1024): md_embed_fused, ONE launch = first embedding + 3 message-passing iterations of all L
    layers; algorithmic bytes per launch = 3 * sum_l (4(n+1) + 4 nnz_l + 8 D n), i.e. the SURVEY 8(d) compulsory floor
    of the three layer-iterations it replaces (H tiles of N<=1000 graphs are cache-resident, so the floor is the figure
    of record), n = active nodes, nnz_l = alive directed edges.  Unfused path: one md_embed_iter launch = one iteration
    of all layers, bytes = sum_l floor."""
    import ctypes as C
    from mdcommunity_b200 import _lib
    ro = mdb.BatchRollout(batch, w)
    lib = ro.lib
    evs, byts = [], []
    env_evs, env_byts = [], []
    D = 64
    N_nodes = torch.as_tensor([g.num_nodes for g in batch.graphs], dtype=torch.float64)
    E_static = batch.n_edges.cpu().double()  # [L, B] undirected edges
    while ro.num_done() < batch.B and len(evs) < 3000:
        n = int(ro.n_act.sum())
        nnz = 2 * ro.alive_edges.sum(1).cpu().numpy()
        live = (ro.done == 0).cpu().double()
        floor = sum(4 * (n + 1) + 4 * int(nnz[l]) + 8 * D * n for l in range(batch.L))
        g, s, wp, k, st = ro._g(), C.byref(ro.s), ro.w.ptr(), C.byref(ro.k), ro._stream()
        if ro.fused:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(lib.md_embed_fused(g, s, wp, k, st), "fused")
            e1.record()
            evs.append((e0, e1))
            byts.append(3 * floor)
        else:
            _lib.check(lib.md_embed_init(g, s, wp, k, st), "init")
            for it in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                _lib.check(lib.md_embed_iter(g, s, wp, k, it, st), "iter")
                e1.record()
                evs.append((e0, e1))
                byts.append(floor)
        _lib.check(lib.md_graph_embed(g, s, wp, k, st), "gembed")
        _lib.check(lib.md_qhead(g, s, wp, k, st), "qhead")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ro.step()
        e1.record()
        env_evs.append((e0, e1))
        # SURVEY 8(d), whole-rollout small graphs: per graph-step floor L(4(N+1) + 4 nnz0) + 4N over the live graphs
        env_byts.append(float((live * (4 * N_nodes + sum(4 * (N_nodes + 1) + 8 * E_static[l] for l in range(batch.L)))).sum()))
    torch.cuda.synchronize()
    env_ms = [a.elapsed_time(b) for a, b in env_evs]
    ms = [a.elapsed_time(b) for a, b in evs]
    tot_ms, tot_b = float(sum(ms)), float(sum(byts))
    peak, how = measured_peak()
    ach = tot_b / (tot_ms * 1e-3) / 1e9 if tot_ms > 0 else 0.0
    traffic = traffic_detail = None
    tpath = os.path.join(ROOT, "profiles", "fused_dram_traffic.json")
    if ro.fused and os.path.exists(tpath):  # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture
        traffic_detail = json.load(open(tpath))
        traffic = traffic_detail.get("bytes_per_launch")
    name = ("embed_fused_kernel (first embedding + 3 x [SpMM aggregation + dense transforms + normalise], one launch)"
            if ro.fused else "embed_iter_kernel (fused SpMM aggregation + dense transforms + normalise)")
    model = ("sum over launches of 3 * sum_l (4(n+1)+4nnz_l+8*64*n)" if ro.fused else "sum over launches of sum_l (4(n+1)+4nnz_l+8*64*n)")
    # the first (intact-graph) launches are the heaviest: report them too
    env_tot = float(sum(env_ms))
    env_ach = float(sum(env_byts)) / (env_tot * 1e-3) / 1e9 if env_tot > 0 else 0.0
    lmcc = {"bound": "latency (shared-memory union-find, one CTA per graph); reported against HBM as the north star asks",
            "kernel": "env_step_small (node removal + LMCC recompute + AUDC + status counters)", "achieved": env_ach,
            "peak": peak, "unit": "GB/s", "frac": env_ach / peak, "launches": len(env_ms),
            "avg_launch_ms": env_tot / max(1, len(env_ms)), "first_launch_ms": env_ms[0] if env_ms else None,
            "algorithmic_bytes_model": "sum over launches and live graphs of L*(4(N+1)+4*nnz0_l)+4N (static CSR + state)"}
    return {"bound": "hbm", "kernel": name, "lmcc": lmcc,
            "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
            "traffic_unit": "bytes per launch (first, intact-state launch; compare with first_launch_algorithmic_bytes)",
            "traffic_detail": traffic_detail, "first_launch_algorithmic_bytes": byts[0] if byts else None,
            "peak_source": how, "launches": len(ms), "avg_launch_ms": tot_ms / max(1, len(ms)),
            "first_launch_ms": ms[0] if ms else None,
            "first_launch_gbs": byts[0] / (ms[0] * 1e-3) / 1e9 if ms else None,
            "algorithmic_bytes_model": model + ", n=active nodes, nnz=alive directed edges"}


def c4_leg(dev, steps=20):
    """Second part of BASELINE.json's metric: removal steps/sec on one 3-layer graph of N = 1M (config C4), timed with CUDA
    events per stage; aggregation = the 9 layer-iterations against SURVEY 8(d)'s no-reuse byte model (H_l = 256 MB > L2)."""
    import ctypes as C
    import mdcommunity_b200 as mdb
    from mdcommunity_b200 import _lib, gmm
    N, L, E, D = 1_000_000, 3, 4_000_000, 64
    t0 = time.time()
    g = mdb.Graph_test(*gmm.random_sparse_layers(N, L, E, seed=0), num_nodes=N, cost="unit")
    batch = mdb.GraphBatch([g], dev, "unit")
    w = mdb.QWeights(load_state_dict("unit"), dev)
    ro = mdb.BatchRollout(batch, w, max_steps=steps + 16)
    torch.cuda.synchronize()
    setup_s = time.time() - t0
    lib = ro.lib
    gp, sp, wp, kp, st = ro._g(), C.byref(ro.s), ro.w.ptr(), C.byref(ro.k), ro._stream()
    rows = []
    for step in range(steps + 3):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        n_act, nnz = int(ro.n_act[0]), int(2 * ro.alive_edges[:, 0].sum())
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
        if step >= 3:
            rows.append([e[i].elapsed_time(e[i + 1]) for i in range(4)] + [n_act, nnz])
        if int(ro.done[0]):
            break
    r = np.array(rows, dtype=np.float64)
    med = np.median(r[:, :4], 0)
    step_ms = float(np.median(r[:, :4].sum(1)))
    n_act, nnz = r[0, 4], r[0, 5]
    agg_bytes = 3 * (L * 4 * (n_act + 1) + 4 * nnz + 4 * D * nnz + L * 4 * D * n_act)
    peak, src = measured_peak()
    return {"workload": "C4: one synthetic 3-layer multiplex, N=1000000, 4000000 undirected edges per layer, unit cost, %d greedy steps "
                        "from the intact graph (score all nodes + arg-max + removal + LMCC recompute per step)" % len(r),
            "removal_steps_per_s": 1e3 / step_ms, "ms_per_step": step_ms,
            "median_ms": {"embed_init": med[0], "aggregation_9_layer_iterations": med[1], "graph_embed+qhead": med[2],
                          "env_step_lmcc": med[3]},
            "aggregation_roofline": {"bound": "hbm", "achieved": agg_bytes / (med[1] * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                     "frac": agg_bytes / (med[1] * 1e-3) / 1e9 / peak, "peak_source": src,
                                     "algorithmic_bytes_model": "9 x (4(n+1) + 4 nnz + 256 nnz + 256 n): SURVEY 8(d) no-reuse model; "
                                                                "iteration 0 reads alive degrees instead of rows (see DESIGN 6.13)"},
            "timing": "CUDA events per stage, median over the steps, 3 warm-up steps", "setup_s": setup_s,
            "lmcc_rank_after": int(ro.rank[0])}


def run_ours(args, cost, n, per_gpu, desc):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "INFO"):
            os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line the driver parses
        dist.init_process_group("nccl", device_id=dev)
    import mdcommunity_b200 as mdb

    graphs = make_graphs(cost, n, per_gpu, rank * per_gpu)
    w = mdb.QWeights(load_state_dict(cost), dev)
    host = mdb.GraphBatch.prepare_host(graphs, cost)
    batch = mdb.GraphBatch(graphs, dev, cost, host=host)
    ro = mdb.BatchRollout(batch, w)
    T = ro.max_steps
    rec_w = 2 + 2 * T  # per-graph record: audc f64, n_steps, then actions / ranks
    gather_buf = torch.empty(world * per_gpu, rec_w, dtype=torch.float64, device=dev) if world > 1 else None

    def one_step():
        ro.reset()
        ro.run()
        if world > 1:  # the only collective of the path: gather per-graph AUDC + removal sequences
            rec = torch.cat([ro.score[:, None], ro.n_steps[:, None].double(), ro.trace_action.double(),
                             ro.trace_rank.double()], 1)
            dist.all_gather_into_tensor(gather_buf, rec)

    for _ in range(args.warmup):
        one_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = ro.kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        one_step()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms[0])
    launches = ro.kernel_launches - launches0
    res = ro.results()
    mean_steps = float(res["n_steps"].mean())
    mean_audc = float(res["score"].mean())

    # ---- end-to-end through the public API with host buffers (H2D of the CSR, D2H of the records) ------------
    def e2e_step():
        b = mdb.GraphBatch(graphs, dev, cost, host=host)  # H2D from pinned memory
        r = mdb.BatchRollout(b, w)
        out = r.run()  # D2H: score, n_steps, actions, ranks, maxcc
        if world > 1:
            rec = torch.cat([r.score[:, None], r.n_steps[:, None].double(), r.trace_action.double(),
                             r.trace_rank.double()], 1)
            dist.all_gather_into_tensor(gather_buf, rec)
        return b, out

    for _ in range(args.warmup):  # the allocator has to see the two-batches-alive pattern before the clock starts
        e2e_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        b, out = e2e_step()
    torch.cuda.synchronize()
    e2e_ms = torch.tensor([1e3 * (time.perf_counter() - t0)], device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_ms = float(e2e_ms[0])
    d2h = sum(int(v.nbytes) for v in out.values())
    clocks = sampler.stop() if sampler else None

    if rank == 0:
        roof = roofline_pass(mdb, batch, w)
        total_graphs = world * per_gpu * args.steps
        line = {
            "metric": "dismantled graphs/sec", "value": total_graphs / (ms * 1e-3), "unit": "graphs/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "graphs_per_gpu": per_gpu, "num_nodes": n, "layers": 2, "cost": cost,
                       "weights": "reference checkpoint nrange_30_50_iter_100000 (tests/golden/weights_%s.npz)" % cost,
                       "generator": "mdcommunity_b200.gmm (vectorised GMM, seeds rank*B+i)",
                       "l2": "no flush: per-step working set (4 embedding buffers, %.0f MB) exceeds the 126 MB L2"
                             % (4 * batch.n_total * 256 / 1e6),
                       "mean_removals_per_graph": mean_steps, "mean_audc": mean_audc,
                       "sharding": "independent graphs per rank, NCCL all_gather of per-graph records"},
            "e2e": {"value": total_graphs / (e2e_ms * 1e-3), "unit": "graphs/s", "h2d_bytes_per_step": int(batch.h2d_bytes),
                    "d2h_bytes_per_step": int(d2h), "timing": "wall clock, max over ranks"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof,
            "removal_steps_per_s": world * per_gpu * mean_steps * args.steps / (ms * 1e-3),
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            torch.set_num_threads(min(16, cores))
            k = 4 if n >= 1000 else 16
            dt, st = oracle_rollouts(cost, graphs[:k])
            line["cpu_baseline"] = {"value": k / dt, "unit": "graphs/s", "cores": torch.get_num_threads(), "kind": "port",
                                    "host_cores": cores, "steps_per_s": st / dt,
                                    "sample": "first %d graphs of the GPU batch, oracle/md_oracle.py rollout, 1 process" % k}
        if world == 1 and args.workload == "c2" and not args.no_c4:
            torch.cuda.empty_cache()
            try:
                line["c4"] = c4_leg(dev)
            except Exception as exc:  # the headline line must survive a failure of the secondary leg
                line["c4"] = {"error": repr(exc)}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--graphs-per-gpu", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c4", action="store_true", help="skip the N = 1M removal-steps/s leg (about 30 s)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3  # timing rule: at least 3 warm-up steps
    cost, n, per_gpu, desc = WORKLOADS[args.workload]
    if args.graphs_per_gpu:
        per_gpu = args.graphs_per_gpu
    if args.impl == "reference":
        run_reference(args, cost, n, desc)
    else:
        run_ours(args, cost, n, per_gpu, desc)


if __name__ == "__main__":
    main()
