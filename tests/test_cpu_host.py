"""CPU-only checks: the C-ABI library loads and exports every symbol the header declares, host-side graph logic,
the synthetic generator, and the graph-sharding gather on a 2-process gloo group.  No compute calls (no GPU here)."""
import ctypes
import os
import re
import socket
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_library_builds_and_exports_header_symbols():
    from mdcommunity_b200 import _lib, build
    path = build.build()
    assert os.path.exists(path)
    h = ctypes.CDLL(path)
    header = open(os.path.join(ROOT, "include", "mdcommunity_b200.h")).read()
    declared = set(re.findall(r"\b(md_[a-z0-9_]+)\s*\(", header))
    declared -= {"md_stream_t"}
    assert {"md_env_reset", "md_env_step", "md_embed_iter", "md_qhead", "md_rollout_steps", "md_spmm_csr_f32"} <= declared
    for sym in declared:
        assert hasattr(h, sym), "library does not export %s" % sym
    for sym in _lib.exported_symbols():
        assert hasattr(h, sym)
        assert sym in declared, "%s is bound in Python but not declared in the header" % sym
    h.md_version.restype = ctypes.c_char_p
    assert b"sm_100a" in h.md_version()
    h.md_env_small_max_stride.restype = ctypes.c_int32
    assert h.md_env_small_max_stride() >= 1024


def test_struct_layouts_match_compiled_abi():
    """The ctypes mirrors must have exactly the sizes the compiled library reports (checked again at load time)."""
    from mdcommunity_b200 import _lib
    h = _lib.lib()
    sizes = (ctypes.c_int32 * 4)()
    h.md_abi_sizes(sizes)
    assert list(sizes) == [ctypes.sizeof(_lib.MdGraphs), ctypes.sizeof(_lib.MdState), ctypes.sizeof(_lib.MdWeights),
                           ctypes.sizeof(_lib.MdWork)]


def test_product_path_refuses_to_run_without_cuda():
    import mdcommunity_b200 as mdb
    g = mdb.Graph_test(np.array([[0, 1], [1, 2]]), np.array([[0, 2], [1, 2]]), num_nodes=3)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
            mdb.GraphBatch([g])
        with pytest.raises(RuntimeError):
            mdb.MultiDismantler()


def test_csr_row_order_is_edge_list_order():
    from mdcommunity_b200.graph import build_csr
    edges = np.array([[3, 1], [0, 1], [1, 2], [0, 3]])
    rp, col = build_csr(4, edges)
    assert list(rp) == [0, 2, 5, 6, 8]
    assert list(col[rp[1]:rp[2]]) == [3, 0, 2]  # neighbours of node 1 in edge-list order
    assert list(col[rp[0]:rp[1]]) == [1, 3]
    assert list(col[rp[3]:rp[4]]) == [1, 0]


def test_prepare_host_block_diagonal():
    import mdcommunity_b200 as mdb
    g1 = mdb.Graph_test(np.array([[0, 1], [1, 2]]), np.array([[0, 2]]), num_nodes=3, cost="degree")
    g2 = mdb.Graph_test(np.array([[0, 4]]), np.array([[1, 2], [2, 3]]), num_nodes=5, cost="degree")
    h = mdb.GraphBatch.prepare_host([g1, g2], "degree")
    S = h["S"]
    assert S == 128 and h["B"] == 2 and h["L"] == 2
    rp0 = h["rowptr"][0].numpy()
    assert rp0[S] == 4 and rp0[2 * S] == 6 and rp0[3] == 4 and rp0[S - 1] == 4
    col0 = h["col"][0].numpy()
    assert sorted(col0[4:6]) == [S + 0, S + 4]
    assert h["max_graph_nnz"] == 6
    w = h["node_w"][0].numpy()
    assert w[1] == 1.0 and w[0] == 0.5 and w[S + 4] == 1.0
    c = h["node_cost"].numpy()
    assert abs(c[:3].sum() - 2.0) < 1e-12  # sum_v (w0/W0 + w1/W1) = 2 per graph
    with pytest.raises(ValueError):
        mdb.Graph_test(np.array([[0, 0]]), np.array([[0, 1]]), num_nodes=2)
    with pytest.raises(ValueError):
        mdb.Graph_test(np.array([[0, 5]]), np.array([[0, 1]]), num_nodes=2)


def test_gmm_generator_shapes_and_oracle_rollout():
    from mdcommunity_b200 import gmm
    from oracle import md_oracle as orc
    g = gmm.gmm_graph(80, 5, cost="degree")
    assert g.num_nodes == 80 and g.num_layers == 2
    for e in g.edge_list:
        assert e.shape[1] == 2 and np.all(e[:, 0] < e[:, 1]) and len(np.unique(e, axis=0)) == len(e)
    assert abs(max(g.weights[0]) - 1.0) < 1e-15
    og = orc.MultiplexGraph(g.num_nodes, g.edge_list, weights=g.weights)
    env = orc.OracleEnv("degree")
    env.s0(og)
    active, degs, aux = env.status()
    for a in np.nonzero(active)[0][:5]:
        env.stepWithoutReward(int(a))
    assert env.score > 0 and env.MaxCCList[-1] <= 1.0
    lay = gmm.random_sparse_layers(2000, 3, 8000, seed=1)
    assert len(lay) == 3 and all(len(e) == 8000 and np.all(e[:, 0] < e[:, 1]) for e in lay)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_items, width, q):
    import torch.distributed as dist
    from mdcommunity_b200 import sharding
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    mine = sharding.shard_indices(n_items, rank, world)
    # fake per-graph results: everything is a function of the global graph index
    score = [1.0 / (7 + i) for i in mine]
    steps = [i % width for i in mine]
    actions = np.array([[(i * 3 + t) % 50 if t < i % width else -1 for t in range(width)] for i in mine]).reshape(len(mine), width)
    ranks = np.array([[max(0, 40 - t - i % 3) if t < i % width else 0 for t in range(width)] for i in mine]).reshape(len(mine), width)
    rec = sharding.pack_records(score, steps, actions, ranks, width)
    per = (n_items + world - 1) // world
    full = sharding.gather_records(rec, n_items, per)
    sc, st, ac, rk = sharding.unpack_records(full, width)
    ok = all(sc[i] == 1.0 / (7 + i) and st[i] == i % width for i in range(n_items))
    ok = ok and all(ac[i, t] == ((i * 3 + t) % 50 if t < i % width else -1) for i in range(n_items) for t in range(width))
    ok = ok and all(rk[i, t] == (max(0, 40 - t - i % 3) if t < i % width else 0) for i in range(n_items) for t in range(width))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_graph_sharding_gather_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 11, 6, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_bench_reference_arm_line():
    """`bench.py --impl reference` prints one JSON line with the contract's keys (tiny workload)."""
    import json
    import subprocess
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "graphs/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0


def _allreduce_worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import importlib.util
    spec = importlib.util.spec_from_file_location("md_train_host", os.path.join(ROOT, "mdcommunity_b200", "train.py"),
                                                  submodule_search_locations=None)
    # allreduce_gradients is pure torch.distributed host logic: exercise it without the CUDA-only package import
    src = open(os.path.join(ROOT, "mdcommunity_b200", "train.py")).read()
    ns = {"torch": torch, "dist": dist}
    start = src.index("def allreduce_gradients(params):")
    end = src.index("# ----", start)
    exec(src[start:end], ns)
    ps = [torch.nn.Parameter(torch.zeros(3, 2)), torch.nn.Parameter(torch.zeros(5))]
    ps[0].grad = torch.full((3, 2), float(rank + 1))
    ps[1].grad = torch.arange(5, dtype=torch.float32) * (rank + 1)
    flat = ns["allreduce_gradients"](ps)
    if rank == 0:
        out.put((flat.tolist(), ps[0].grad.tolist(), ps[1].grad.tolist()))
    dist.destroy_process_group()


def test_gradient_allreduce_gloo_world2():
    """Data-parallel training step: ONE flat gradient buffer, SUM all-reduce (world 2, gloo)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29600 + os.getpid() % 200
    procs = [ctx.Process(target=_allreduce_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    flat, g0, g1 = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert flat == [3.0] * 6 + [0.0, 3.0, 6.0, 9.0, 12.0]
    assert g0 == [[3.0, 3.0]] * 3 and g1 == [0.0, 3.0, 6.0, 9.0, 12.0]


def test_dense_adjacency_loader_matches_networkx_edge_order(tmp_path):
    """load_npy_pair (the .npy synthetic test sets of testSynthetic, U/MultiDismantler_torch.py:575-587): the edge
    list must come out in the order of nx.from_numpy_array(adj) -> G.edges(), which fixes the neighbour summation order."""
    import networkx as nx
    from mdcommunity_b200 import graph as mg
    rng = np.random.default_rng(0)
    n = 40
    adjs = []
    for k in range(2):
        a = (rng.random((n, n)) < 0.12).astype(np.float64)
        a = np.triu(a, 1)
        a = a + a.T
        adjs.append(a)
        np.save(tmp_path / ("adj%d_0.npy" % (k + 1)), a)
    g = mg.load_npy_pair(str(tmp_path / "adj1_0.npy"), str(tmp_path / "adj2_0.npy"))
    assert g.num_nodes == n and g.num_layers == 2
    for l in range(2):
        ref = np.array(list(nx.from_numpy_array(adjs[l]).edges()), dtype=np.int64).reshape(-1, 2)
        assert np.array_equal(g.edge_list[l], ref)
    with pytest.raises(ValueError):
        mg.edges_from_dense(np.zeros((3, 4)))
