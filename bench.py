#!/usr/bin/env python
"""Benchmark of the greedy-dismantling rollout (BASELINE.json metric: dismantled graphs/sec, N=1000, 2-layer).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c1|c3]

A "step" is one complete lock-step greedy rollout (s0 -> every graph terminal) of one batch of synthetic graphs.
Workload at N=1 (BASELINE.json configs[1], "c2"): degree-cost rollout, 256 two-layer GMM graphs of N=1000 per GPU.
With --gpus N (torchrun, one rank per GPU) every rank owns its own 256 graphs (weak scaling, no data-path
collective); the per-graph records are all-gathered over NCCL at the end of every step inside the timed region.

One JSON line is printed by rank 0 (see the contract in the task statement); extra keys: roofline (dominant C2 kernel:
HBM model + shared-memory model, and the LMCC kernel), cpu_baseline, c4 (removal steps/s at N = 1M with per-kernel
rooflines and a scipy CPU baseline), c3_strong (1024 unit-cost graphs of the three testSynthetic families split over the
ranks: strong scaling), per_rank_ms / gather_ms (multi-GPU).
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


# C3 (BASELINE.json configs[2]): 1024 unit-cost graphs, N = 1000, the three families of testSynthetic
# (U/testSynthetic.py:20 data_g / data_gamma / data_k; knobs of U/GMM.py:10-25), seed = graph index
C3_TOTAL = 1024
C3_G = (0.0, 0.25, 0.5, 0.75, 1.0)
C3_GAMMA = (2.1, 2.5, 3.0)
C3_KBAR = (4.0, 6.0, 8.0, 10.0)


def c3_params(i):
    fam, j = i % 3, i // 3
    g, gamma, kbar = 0.5, 2.5, 6.0
    if fam == 0:
        g = C3_G[j % len(C3_G)]
    elif fam == 1:
        gamma = C3_GAMMA[j % len(C3_GAMMA)]
    else:
        kbar = C3_KBAR[j % len(C3_KBAR)]
    return dict(g=g, gamma=(gamma, gamma), kbar=(kbar, kbar))


def _c3_graph(i):
    from mdcommunity_b200 import gmm
    return gmm.gmm_graph(1000, 50_000 + i, cost="unit", **c3_params(i))


def make_c3_graphs(indices):
    """Graph i of the C3 set for every i in `indices` (host generation, outside every timed region; a process pool when
    the shard is large -- must run before CUDA is initialised in this process)."""
    indices = list(indices)
    if len(indices) > 160 and (os.cpu_count() or 1) >= 4 and not torch.cuda.is_initialized():
        import multiprocessing as mp
        with mp.get_context("fork").Pool(min(16, os.cpu_count())) as pool:
            return pool.map(_c3_graph, indices, chunksize=8)
    return [_c3_graph(i) for i in indices]


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
    evs, byts, smem_b, smem_min = [], [], [], []
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
            fb, mb = smem_model_bytes(n, [int(z) for z in nnz], iters=2 if batch.cost == "unit" else 3)
            smem_b.append(fb)
            smem_min.append(mb)
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
    smem = None
    if ro.fused and smem_b:
        props = torch.cuda.get_device_properties(batch.device)
        try:
            import pynvml
            pynvml.nvmlInit()
            mhz = pynvml.nvmlDeviceGetMaxClockInfo(pynvml.nvmlDeviceGetHandleByIndex(batch.device.index or 0), pynvml.NVML_CLOCK_SM)
        except Exception:
            mhz = 1965
        speak = 128.0 * props.multi_processor_count * mhz * 1e6 / 1e9  # 128 B / clk / SM
        sach = float(sum(smem_b)) / (tot_ms * 1e-3) / 1e9
        smin = float(sum(smem_min)) / (tot_ms * 1e-3) / 1e9
        smem = {"bound": "shared memory bandwidth (the embeddings of a (graph, layer) unit stay resident in one SM's shared "
                         "memory: DRAM traffic is ~7x below the HBM model, so HBM is not what limits this kernel)",
                "achieved": sach, "peak": speak, "unit": "GB/s", "frac": sach / speak,
                "first_launch_gbs": smem_b[0] / (ms[0] * 1e-3) / 1e9, "first_launch_frac": smem_b[0] / (ms[0] * 1e-3) / 1e9 / speak,
                "min_model_gbs": smin, "min_model_frac": smin / speak,
                "peak_source": "128 B/clk/SM x %d SMs x %d MHz (max SM clock)" % (props.multi_processor_count, mhz),
                "model": "traffic of the formulation, per launch and layer-iteration: 260 nnz_l (half rows gathered in both "
                         "passes + 2-byte neighbour codes) + 2432 n (own rows 256, fp16 operand rows staged 512, MMA operand "
                         "fetches 1152, epilogue 512); min_model = 256 nnz_l + 1024 n, what any shared-memory-resident "
                         "formulation moves (the round-1 figure)",
                "ncu": "profiles/r2z_ncu_embed_fused_16warps_summary.txt: 46.5 M LSU wavefronts (5.96 GB) + ~1.8 GB of MMA "
                       "operand fetches in 402 us = 19 TB/s at the intact state, 5.8 M wavefronts are bank conflicts"}
    return {"bound": "hbm", "kernel": name, "lmcc": lmcc, "smem": smem,
            "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
            "traffic_unit": "bytes per launch (first, intact-state launch; compare with first_launch_algorithmic_bytes)",
            "traffic_detail": traffic_detail, "first_launch_algorithmic_bytes": byts[0] if byts else None,
            "peak_source": how, "launches": len(ms), "avg_launch_ms": tot_ms / max(1, len(ms)),
            "first_launch_ms": ms[0] if ms else None,
            "first_launch_gbs": byts[0] / (ms[0] * 1e-3) / 1e9 if ms else None,
            "algorithmic_bytes_model": model + ", n=active nodes, nnz=alive directed edges"}


def c4_cpu_baseline(layers, N, D=64):
    """SURVEY 8(d) / BASELINE.md section 3: C4 cannot run in the reference (dense [n, n] adjacency, L hard-wired to 2), so
    the CPU figures are labelled RESTATEMENTS timed on the host: one neighbour aggregation = scipy.sparse CSR @ dense
    [N, 64] f32, one LMCC recompute = the connected_components-based mutual-component loop of oracle/md_oracle.py after
    removing the node of highest layer-0 degree."""
    import scipy.sparse as sp
    from oracle import md_oracle as orc
    L = len(layers)
    e = layers[0]
    a = sp.csr_matrix((np.ones(2 * len(e), np.float32), (np.concatenate([e[:, 0], e[:, 1]]), np.concatenate([e[:, 1], e[:, 0]]))),
                      shape=(N, N))
    h = np.random.default_rng(0).standard_normal((N, D), dtype=np.float32)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        a @ h
        ts.append(time.perf_counter() - t0)
    agg_s = min(ts)
    og = orc.MultiplexGraph(N, layers, max_rank=N)
    removed = np.zeros(N, bool)
    t0 = time.perf_counter()
    lab = orc.mutual_labels(og, removed, np.zeros(N, np.int64))
    s0_s = time.perf_counter() - t0
    removed[int(np.argmax(np.bincount(e.ravel(), minlength=N)))] = True
    t0 = time.perf_counter()
    orc.mutual_labels(og, removed, lab)
    mcc_s = time.perf_counter() - t0
    step_s = 3 * L * agg_s + mcc_s  # the 9 aggregations + one LMCC recompute of a step (dense transforms / Q head not counted)
    return {"kind": "port", "cores": 1, "value": 1.0 / step_s, "unit": "removal steps/s",
            "aggregation_ms": 1e3 * agg_s, "lmcc_recompute_ms": 1e3 * mcc_s, "lmcc_s0_ms": 1e3 * s0_s,
            "sample": "one scipy.sparse CSR[1M x 1M, 8M nnz] @ [1M, 64] f32 product (best of 3) and one mutual-component "
                      "recompute (scipy connected_components loop) on the C4 graph; step = 9 aggregations + 1 recompute; "
                      "labelled restatements (the reference cannot run this configuration), single-threaded scipy"}


def c4_leg(dev, steps=100):
    """Second part of BASELINE.json's metric: removal steps/sec on one 3-layer graph of N = 1M (config C4), >= 100
    consecutive greedy steps from the intact graph, every kernel of the step timed with CUDA events on its stream and
    set against its own algorithmic bytes (H_l = 256 MB > L2, so the no-reuse model is the figure of record
    for the aggregation)."""
    import ctypes as C
    import mdcommunity_b200 as mdb
    from mdcommunity_b200 import _lib, gmm
    N, L, E, D = 1_000_000, 3, 4_000_000, 64
    t0 = time.time()
    layers = gmm.random_sparse_layers(N, L, E, seed=0)
    g = mdb.Graph_test(*layers, num_nodes=N, cost="unit")
    batch = mdb.GraphBatch([g], dev, "unit")
    w = mdb.QWeights(load_state_dict("unit"), dev)
    ro = mdb.BatchRollout(batch, w, max_steps=steps + 16)
    torch.cuda.synchronize()
    setup_s = time.time() - t0
    lib = ro.lib
    gp, sp, wp, kp, st = ro._g(), C.byref(ro.s), ro.w.ptr(), C.byref(ro.k), ro._stream()
    names = ["embed_init", "agg0", "xform0", "agg1", "xform1", "agg2", "xform2", "graph_embed", "qhead", "env_step"]
    rows, shapes, rounds = [], [], []
    WARM = 3
    for step in range(steps + WARM):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
        n_act, nnz = int(ro.n_act[0]), int(2 * ro.alive_edges[:, 0].sum())
        k = 0
        e[k].record()
        _lib.check(lib.md_embed_init(gp, sp, wp, kp, st), "init")
        k += 1
        e[k].record()
        for it in range(3):  # the two kernels of an iteration, enqueued one by one (md_debug_set_flags bits 14 / 15)
            lib.md_debug_set_flags(16384)
            _lib.check(lib.md_embed_iter(gp, sp, wp, kp, it, st), "agg")
            k += 1
            e[k].record()
            lib.md_debug_set_flags(32768)
            _lib.check(lib.md_embed_iter(gp, sp, wp, kp, it, st), "xform")
            k += 1
            e[k].record()
        lib.md_debug_set_flags(0)
        _lib.check(lib.md_graph_embed(gp, sp, wp, kp, st), "gembed")
        k += 1
        e[k].record()
        _lib.check(lib.md_qhead(gp, sp, wp, kp, st), "qhead")
        k += 1
        e[k].record()
        ro.step()
        k += 1
        e[k].record()
        torch.cuda.synchronize()
        if step >= WARM:
            rows.append([e[i].elapsed_time(e[i + 1]) for i in range(len(names))])
            shapes.append((n_act, nnz))
            rounds.append(int(ro.flags[7]) >> 24)  # prune rounds of this step's LMCC recompute (kernel-side counter)
        if int(ro.done[0]):
            break
    r = np.array(rows, dtype=np.float64)
    med = dict(zip(names, np.median(r, 0)))
    step_ms = float(np.median(r.sum(1)))
    n_act = float(np.median([x[0] for x in shapes]))
    nnz = float(np.median([x[1] for x in shapes]))  # directed alive entries, all layers
    R = float(np.median(rounds))
    peak, src = measured_peak()

    def roof(ms, byts, **kw):
        a = byts / (ms * 1e-3) / 1e9
        return dict({"bound": "hbm", "ms": ms, "algorithmic_bytes": byts, "achieved": a, "peak": peak, "unit": "GB/s", "frac": a / peak}, **kw)

    b_rows = L * 4 * (n_act + 1) + 4 * nnz + 4 * D * nnz + L * 4 * D * n_act  # one md_agg launch (all layers), no-reuse
    b_deg = L * 4 * (n_act + 1) + 4 * nnz + 4 * nnz + L * 4 * D * n_act         # iteration 0: 4-byte degree gather per entry
    b_xf = L * 3 * 4 * D * n_act                                                # read P, read H, write H'
    b_xf0 = L * (2 * 4 * D * n_act + 4 * n_act)                                 # iteration 0: H0 comes from the degrees
    b_qh = L * 4 * D * n_act + 4 * n_act + 4 * n_act                            # H_l rows, ids, q
    b_sweep = 4 * (n_act + 1) + 2 * 4 * nnz / L + 4 * n_act + n_act / 8         # one edge sweep of one layer
    # env_step_large: the mutual-component update is INCREMENTAL (multi-source searches from the end points of the edges that
    # died: R search tasks per step visiting ~1e4 of the 1e6 nodes; SURVEY 8(d): incremental variants count only the rows
    # visited, which is negligible here), so the step's algorithmic bytes are the two remaining full passes per layer:
    # alive degrees / component sizes, and the emission of the alive CSR in compact order.
    sweeps = 2 * L
    kern = {
        # unit cost: the embedding after iteration 0 is a function of the alive degree (md_agg.cu, agg_deg_kernel<1>), so
        # iteration 0 launches nothing (agg0 / xform0 time an empty call) and iteration 1 gathers 4-byte degrees
        "agg_deg<1> (iteration 1: degree gather, table H1(d))": roof(med["agg1"], b_deg),
        "agg_ring (iteration 2: 256-byte row gather)": roof(med["agg2"], b_rows),
        "embed_iter_tc<PRE> (transform, iteration 2)": roof(med["xform2"], b_xf),
        "embed_iter_tc<PRE> (transform, iteration 1, own rows from the degrees)": roof(med["xform1"], b_xf0),
        "qhead_tc<3>": roof(med["qhead"], b_qh),
        "env_step_large": roof(med["env_step"], sweeps * b_sweep, search_tasks=R, L=L, sweeps=sweeps, B_sweep=b_sweep,
                               model="2L * B_sweep: per layer one sweep for alive degrees / component sizes and one for the "
                                     "alive-CSR emission; the incremental mutual-component update (search_tasks multi-source "
                                     "searches per step) visits ~1e4 nodes and is not counted"),
    }
    agg_ms = sum(med[k] for k in names[1:7])
    agg_bytes = b_deg + b_rows + b_xf0 + b_xf
    step_bytes = agg_bytes + b_qh + sweeps * b_sweep
    out = {"workload": "C4: one synthetic 3-layer multiplex, N=1000000, 4000000 undirected edges per layer, unit cost, %d consecutive "
                       "greedy steps from the intact graph (score all nodes + arg-max + removal + LMCC recompute per step)" % len(r),
           "removal_steps_per_s": 1e3 / step_ms, "ms_per_step": step_ms, "steps_timed": len(r),
           "median_ms": {k: float(v) for k, v in med.items()},
           "stage_ms": {"embedding_iterations": agg_ms, "graph_embed+qhead": med["graph_embed"] + med["qhead"],
                        "env_step_lmcc": med["env_step"]},
           "aggregation_roofline": dict(roof(agg_ms, agg_bytes), peak_source=src,
                                        algorithmic_bytes_model="bytes actually required: iteration 0 = nothing (H1 is a function "
                                        "of the alive degree for unit-cost features); iteration 1 = degree gather (4 B per entry) + "
                                        "transform with own rows from the degrees; iteration 2 = 4(n+1) + 4 nnz + 256 nnz + 256 n "
                                        "(SURVEY 8(d) no-reuse) + 3 x 256 n for the transform"),
           "kernels": kern, "step_roofline": roof(step_ms, step_bytes),
           "timing": "CUDA events around every launch, median over the steps, %d warm-up steps" % WARM, "setup_s": setup_s,
           "lmcc_rank_after": int(ro.rank[0]), "n_act_median": n_act, "alive_directed_entries_median": nnz}
    del ro, batch
    torch.cuda.empty_cache()
    try:
        out["cpu_baseline"] = c4_cpu_baseline(layers, N)
    except Exception as exc:
        out["cpu_baseline"] = {"error": repr(exc)}
    return out


def _nccl_log_setup(rank):
    """NCCL's INFO log goes to a FILE (one per rank) instead of being muted: stdout stays the one JSON line the driver
    parses, and the line can still say how many ranks the communicator really had."""
    d = os.path.join(os.environ.get("GRAFT_REPO_ROOT", ROOT), "gpurun_out")
    try:
        os.makedirs(d, exist_ok=True)
    except OSError:
        d = "/tmp"
    os.environ["NCCL_DEBUG"] = "INFO"
    os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")
    os.environ["NCCL_DEBUG_FILE"] = os.path.join(d, "nccl_bench_rank%d.log" % rank)
    return os.environ["NCCL_DEBUG_FILE"]


def _nccl_log_summary(path):
    import re
    out = {"log": path, "nranks": None, "nvls": None}
    try:
        txt = open(path, errors="replace").read()
    except OSError:
        return out
    m = re.findall(r"nranks (\d+)", txt)
    if m:
        out["nranks"] = max(int(x) for x in m)
    out["nvls"] = bool(re.search(r"NVLS", txt))
    return out


def smem_model_bytes(n, nnz_per_layer, D=64, iters=3):
    """Shared-memory traffic of md_embed_fused's formulation, per launch (`iters` iterations, all layers, two column-half
    passes per iteration), in bytes; n = active rows of the batch, nnz = alive directed entries per layer.  Per layer-iteration:
      gather            256 nnz   one 128-byte half row per entry and pass
      neighbour ids       4 nnz   2-byte row codes, read in both passes
      own rows          256 n     read for the own-embedding operand and the column sums
      staging writes    512 n     [P_hi | P_lo] and [H_hi | H_lo] fp16 operand rows, 2 x 128 B per row and pass
      MMA operand reads 1152 n    12 MMAs per 128-row tile and pass, each fetching a 4 KB A slice and a 2 KB B slice
      epilogue          512 n     new embedding: 128 B into the resident half, 128 B into the staging tile, read by the bulk
                                  copy to the scratch image and written back by the reload for pass 1
    -> (260 nnz + 2432 n) per layer-iteration.  `min` is the part no shared-memory-resident formulation avoids (gather, own
    rows, operand rows read once, new embedding written once: 256 nnz + 1024 n), the figure round 1 reported."""
    full = iters * sum(260 * z + 2432 * n for z in nnz_per_layer)
    mn = iters * sum(4 * D * z + 4 * 4 * D * n for z in nnz_per_layer)
    return full, mn


def c3_strong_leg(mdb, dist, dev, world, rank, graphs, idx, w, steps=3, warmup=1):
    """BASELINE.json configs[2]: 1024 unit-cost graphs (three testSynthetic families) split over the ranks
    (sharding.shard_indices: graph i -> rank i mod W), every rank rolls out its shard, one all_gather of the per-graph
    records (sharding.gather_records) per pass.  STRONG scaling: total work fixed, value = 1024 graphs / max-over-ranks time."""
    from mdcommunity_b200 import sharding
    batch = mdb.GraphBatch(graphs, dev, "unit")
    ro = mdb.BatchRollout(batch, w, max_steps=1000)
    per_rank_max = (C3_TOTAL + world - 1) // world
    width = 200  # removal-sequence prefix carried in the record (the full sequences stay on the owning rank)
    full = None

    def one_pass():
        nonlocal full
        ro.reset()
        ro.run()
        rec = torch.cat([ro.score[:, None], ro.n_steps[:, None].double(), ro.trace_action[:, :width].double(),
                         ro.trace_rank[:, :width].double()], 1)
        full = sharding.gather_records(rec, C3_TOTAL, per_rank_max, device=dev)

    for _ in range(warmup):
        one_pass()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        one_pass()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms[0]) / steps
    score, n_steps, _, _ = sharding.unpack_records(full, width)
    return {"workload": "C3: 1024 unit-cost 2-layer GMM graphs, N=1000; families data_g (g in %s), data_gamma (gamma in %s), "
                        "data_k (kbar in %s), other knobs g=0.5 gamma=2.5 kbar=6, seed 50000+i; graph i -> rank i mod W" % (
                            list(C3_G), list(C3_GAMMA), list(C3_KBAR)),
            "scaling": "strong", "n_gpus": world, "graphs_total": C3_TOTAL, "graphs_per_gpu": len(graphs),
            "value": C3_TOTAL / (ms * 1e-3), "unit": "graphs/s", "ms_per_pass": ms, "passes": steps,
            "mean_removals_per_graph": float(n_steps.mean()), "max_removals": int(n_steps.max()), "mean_audc": float(score.mean()),
            "timing": "CUDA events, reset + rollout + all_gather of the records per pass, max over ranks"}


def run_ours(args, cost, n, per_gpu, desc):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; there is no CPU fallback for the product path")
    # ---- synthetic inputs (host; before CUDA is initialised: the large C3 shard is generated by a process pool) ----
    from mdcommunity_b200 import sharding
    do_c3 = args.workload == "c2" and not args.no_c3
    c3_idx = sharding.shard_indices(C3_TOTAL, rank, world) if do_c3 else []
    c3_graphs = make_c3_graphs(c3_idx) if do_c3 else []
    graphs = make_graphs(cost, n, per_gpu, rank * per_gpu)

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    nccl_log = None
    if world > 1:
        nccl_log = _nccl_log_setup(rank)
        dist.init_process_group("nccl", device_id=dev)
    import mdcommunity_b200 as mdb

    from mdcommunity_b200.rollout import RolloutPipeline
    w = mdb.QWeights(load_state_dict(cost), dev)
    host = mdb.GraphBatch.prepare_host(graphs, cost)
    batch = mdb.GraphBatch(graphs, dev, cost, host=host)
    depth = max(1, args.depth)
    tail = bool(args.tail)
    ros = [mdb.BatchRollout(batch, w, tail=tail) for _ in range(depth)]  # `depth` batches in flight (same graphs, own state)
    ro = ros[0]
    T = ro.max_steps
    gather_ev = []
    pipe = RolloutPipeline(dev, depth=depth)
    timed = [False]

    gathered = [None]
    width = [None]                                        # columns of the trace buffers that are gathered
    overflow = torch.zeros(1, dtype=torch.int32, device=dev)

    pending = []  # packed records of the finished batches of the running call, waiting for the collective

    def flush():
        """ONE all_gather for all the batches a run_steps / e2e_steps call has finished (inside the timed region).  A
        collective per batch couples the ranks at every batch: a rank whose rollouts happen to be short waits -- with NCCL's
        CTAs resident and spinning, on SMs where no embed_fused CTA (the whole register file) can start -- for the rank with
        the longest one, 8 to 10 % of the step at 2 to 8 GPUs; the records of a 20-step job are 16 MB per rank."""
        if world > 1 and pending:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            rec = torch.stack(pending)
            out = torch.empty((world,) + tuple(rec.shape), dtype=torch.float64, device=dev)
            dist.all_gather_into_tensor(out, rec)
            gathered[0] = out
            if timed[0]:
                b.record()
                gather_ev.append((a, b))
        pending.clear()

    def gather(r):
        """The only collective of the path: per-graph AUDC (f64, bit-exact) + n_steps + removal sequence + LMCC trace of the
        finished batches: packed here per batch, gathered by flush().  The traces are [B, N] buffers of which a rollout fills ~50-70 columns: the ranks agree ONCE (first
        call, during warm-up: one 4-byte MAX all-reduce read back by the host) on a record width = the longest sequence + 25 %
        head room; afterwards a batch is handed to the collective without any host round trip (a host read of a collective's
        result per batch stalls the other batches in flight until the slowest rank gets there: 8 GPUs lost 17 % to it).  A
        sequence longer than the width sets a device-side flag that is checked after the timed region."""
        if world > 1:
            mx = r.n_steps.max().reshape(1).clone()
            if width[0] is None:
                dist.all_reduce(mx, op=dist.ReduceOp.MAX)
                width[0] = min(T, (int(1.25 * int(mx.item())) + 16) // 16 * 16)
            used = width[0]
            overflow.add_((mx > used).to(torch.int32))
            rec = torch.cat([r.score[:, None], r.n_steps[:, None].double(), r.trace_action[:, :used].double(),
                             r.trace_rank[:, :used].double()], 1)
            pending.append(rec)

    def run_steps(k):
        """k steps = k complete rollouts of the batch, `depth` of them in flight at a time (RolloutPipeline): a step's
        under-occupied straggler phase overlaps the next step's full-width kernels."""
        t0 = time.perf_counter()
        pipe.run(k, lambda j: ros[j % depth], lambda j, r: gather(r))
        t1 = time.perf_counter()
        flush()
        if os.environ.get("MDB200_BENCH_TRACE"):
            torch.cuda.synchronize()
            sys.stderr.write("rank %d: run_steps(%d): pipeline host time %.2f ms, flush (to sync) %.2f ms\n"
                             % (rank, k, 1e3 * (t1 - t0), 1e3 * (time.perf_counter() - t1)))

    run_steps(max(args.warmup, 2 * depth))  # every rollout object has run twice: its CUDA graph exists
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = sum(r.kernel_launches for r in ros)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    timed[0] = True
    e0.record()
    run_steps(args.steps)
    e1.record()
    torch.cuda.synchronize()
    timed[0] = False
    if world > 1:
        dist.barrier()
    my_ms = e0.elapsed_time(e1)
    ms = torch.tensor([my_ms], device=dev)
    per_rank_ms = [my_ms / args.steps]
    gather_ms = None
    if world > 1:
        if int(overflow.item()):
            raise RuntimeError("a removal sequence was longer than the gathered record width %d" % width[0])
        allms = torch.zeros(world, device=dev)
        dist.all_gather_into_tensor(allms, ms.clone())
        per_rank_ms = [float(x) / args.steps for x in allms.cpu()]
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        gather_ms = float(np.sum([a.elapsed_time(b) for a, b in gather_ev])) / args.steps
    ms = float(ms[0])
    launches = sum(r.kernel_launches for r in ros) - launches0
    res = ro.results()
    mean_steps = float(res["n_steps"].mean())
    mean_audc = float(res["score"].mean())
    # latency of ONE batch with nothing else in flight (reset + rollout), for reference next to the pipelined throughput
    lat = []
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ro.reset()
        ro.run()
        b.record()
        torch.cuda.synchronize()
        lat.append(a.elapsed_time(b))
    single_batch_ms = float(np.median(lat))

    # ---- end-to-end through the public API with host buffers (H2D of the CSR, D2H of the records) ------------
    e2e_out = [None]

    # a pool of `depth` resident batches (device CSR buffers + rollout state + captured CUDA graph); every step streams its
    # input -- the block-diagonal CSR in pinned host memory -- into one of them (GraphBatch.upload: H2D on the job's
    # stream, no allocation), rolls it out and reads the records back
    e2e_pool = [mdb.BatchRollout(mdb.GraphBatch(graphs, dev, cost, host=host), w, tail=tail) for _ in range(depth)]

    def e2e_make(j):
        r = e2e_pool[j % depth]
        r.batch.upload(host)
        return r

    def e2e_finish(j, r):
        e2e_out[0] = r.results()  # D2H: score, n_steps, actions, ranks, maxcc
        gather(r)

    def e2e_steps(k):
        pipe.run(k, e2e_make, e2e_finish, reset=True)
        flush()

    e2e_steps(max(args.warmup, 2 * depth))  # every pooled rollout has run twice: its CUDA graph exists
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    e2e_steps(args.steps)
    torch.cuda.synchronize()
    e2e_ms = torch.tensor([1e3 * (time.perf_counter() - t0)], device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_ms = float(e2e_ms[0])
    out = e2e_out[0]
    d2h = sum(int(v.nbytes) for v in out.values())
    clocks = sampler.stop() if sampler else None
    del out

    c3 = None
    if do_c3:
        try:
            wu = mdb.QWeights(load_state_dict("unit"), dev)
            c3 = c3_strong_leg(mdb, dist, dev, world, rank, c3_graphs, c3_idx, wu)
        except Exception as exc:  # the headline line must survive a failure of a secondary leg
            c3 = {"error": repr(exc)}
            if world > 1:
                raise

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
                       "sharding": "independent graphs per rank; ONE NCCL all_gather of the per-graph records of all timed steps, inside the timed region",
                       "batches_in_flight": depth, "tail_kernel": bool(ro.tail),
                       "step": "one complete rollout (s0 -> all terminal) of the batch; %d steps are in flight at a time on "
                               "their own streams (RolloutPipeline), all of them inside the timed region" % depth},
            "e2e": {"value": total_graphs / (e2e_ms * 1e-3), "unit": "graphs/s", "h2d_bytes_per_step": int(batch.h2d_bytes),
                    "d2h_bytes_per_step": int(d2h), "timing": "wall clock, max over ranks",
                    "path": "per step: GraphBatch.upload (the block-diagonal CSR from pinned host memory into a pooled resident "
                            "batch, H2D on the job's stream) -> s0 -> rollout -> BatchRollout.results() (D2H into pinned buffers) "
                            "-> record gather; %d pooled batches in flight" % depth},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof,
            "removal_steps_per_s": world * per_gpu * mean_steps * args.steps / (ms * 1e-3),
            "per_rank_ms_per_step": per_rank_ms, "gather_span_ms_per_step": gather_ms,
            "gather_note": "ONE all_gather of the records of all timed steps at the end of the timed region (records packed per batch on the batch's stream; width agreed once during warm-up); span of that collective divided by the steps (a collective per finished batch, as in the previous "
                           "builds, coupled the ranks at every batch: 8-10 % of the step at 2-8 GPUs)",
            "single_batch_latency_ms": single_batch_ms,
        }
        if nccl_log:
            line["nccl"] = _nccl_log_summary(nccl_log)
        if c3 is not None:
            line["c3_strong"] = c3
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            torch.set_num_threads(min(16, cores))
            k = min(len(graphs), 4 if n >= 1000 else 16)
            dt, st = oracle_rollouts(cost, graphs[:k])
            line["cpu_baseline"] = {"value": k / dt, "unit": "graphs/s", "cores": torch.get_num_threads(), "kind": "port",
                                    "host_cores": cores, "steps_per_s": st / dt,
                                    "sample": "first %d graph(s) of the GPU batch, oracle/md_oracle.py rollout, 1 process" % k}
        if world == 1 and args.workload == "c2" and not args.no_c4:
            del ro, ros, batch
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
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--graphs-per-gpu", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c4", action="store_true", help="skip the N = 1M removal-steps/s leg (about 30 s)")
    ap.add_argument("--no-c3", action="store_true", help="skip the 1024-graph strong-scaling leg (C3)")
    ap.add_argument("--depth", type=int, default=10, help="batches in flight (RolloutPipeline); 1 = strictly one after the other "
                    "(measured at 20 steps, final round-2 build: depth 6 / 8 / 10 / 12 / 16 = 22.6 / 23.0 / 23.5 / 23.6 / 23.6 k graphs/s)")
    ap.add_argument("--tail", action="store_true", help="hand graphs with <= 64 active nodes over to md_rollout_tail")
    ap.add_argument("--fused-pack", type=int, default=0, help="md_debug_set_fused_pack (A/B): 0 = default bin budget, > 0 fixed, < 0 every unit alone")
    ap.add_argument("--fused-legacy", action="store_true", help="md_embed_fused without packing of small units (debug bit 19; A/B)")
    args = ap.parse_args()
    if args.impl == "ours" and (args.fused_pack or args.fused_legacy):
        from mdcommunity_b200 import _lib
        _lib.lib().md_debug_set_fused_pack(args.fused_pack)
        if args.fused_legacy:
            _lib.lib().md_debug_set_flags(1 << 19)
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
