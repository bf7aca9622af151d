"""Graph containers: host-side multiplex graph (mirror of the reference's graph.Graph_test fields) and the
device-resident batch (multi-layer CSR, built once, never rebuilt during a rollout).

Reference: /root/reference/code/MultiDismantler_unit_cost/graph.py:8-30 (Graph), :70-84 (Graph_test), :53-68 (GSet);
degree-cost node weights /root/reference/code/MultiDismantler_degree_cost/graph.py:39-51.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib

TILE_M = _lib.TILE_M


def _pin(t):
    return t.pin_memory() if torch.cuda.is_available() else t


class Graph_test:
    """Static multiplex graph given as per-layer undirected edge lists (the reference passes two nx.Graph objects;
    any object with .nodes/.edges works, as do plain arrays).

    Fields mirror the reference: num_nodes, edge_list, num_edges, weights, max_rank.  max_rank (LMCC of the intact
    graph, reference ori_rank) is filled in by the GPU when the graph is placed in a GraphBatch.
    """

    def __init__(self, G1, G2=None, *more_layers, num_nodes=None, cost="unit"):
        layers = [G1] + ([G2] if G2 is not None else []) + list(more_layers)
        self.edge_list = []
        for G in layers:
            if hasattr(G, "edges") and hasattr(G, "nodes"):
                if num_nodes is None:
                    num_nodes = len(G.nodes)
                e = np.array(list(G.edges()), dtype=np.int64).reshape(-1, 2)
            else:
                e = np.asarray(G, dtype=np.int64).reshape(-1, 2)
            self.edge_list.append(e)
        if num_nodes is None:
            num_nodes = int(max((int(e.max()) + 1 if e.size else 0) for e in self.edge_list))
        self.num_nodes = int(num_nodes)
        for e in self.edge_list:
            if e.size and (e.min() < 0 or e.max() >= self.num_nodes):
                raise ValueError("edge endpoint out of range")
            if e.size and np.any(e[:, 0] == e[:, 1]):
                raise ValueError("self loops are not allowed (the reference's read_multiplex drops them)")
        self.num_layers = len(self.edge_list)
        self.num_edges = [int(len(e)) for e in self.edge_list]
        self.max_rank = 0
        self.cost = cost
        self.weights = None
        if cost == "degree":
            self.cal_degree()

    def cal_degree(self):
        """D/graph.py:39-51: weights[l][v] = deg_l(v) / max_v deg_l(v) in float64."""
        self.weights = []
        for e in self.edge_list:
            deg = np.bincount(e.ravel(), minlength=self.num_nodes).astype(np.int64)
            self.weights.append(deg / max(int(deg.max()), 1))


def edges_from_dense(adj):
    """Edge list of a dense symmetric adjacency matrix in the order the reference obtains it:
    nx.from_numpy_array(adj) followed by G.edges() (U/MultiDismantler_torch.py:585-587) yields each undirected edge once,
    as (u, v) with u < v, in row-major order of the upper triangle.  The order matters: neighbour rows are summed in
    edge-list order (see build_csr).  Diagonal entries (self loops) are dropped."""
    a = np.asarray(adj)
    if a.ndim != 2 or a.shape[0] != a.shape[1]:
        raise ValueError("adjacency must be a square matrix")
    sym = (a != 0) | (a.T != 0)
    return np.argwhere(np.triu(sym, 1)).astype(np.int64)


def load_npy_pair(path1, path2, cost="unit"):
    """Two dense adjacency .npy files (the synthetic test sets of testSynthetic: adj1_i.npy / adj2_i.npy,
    U/MultiDismantler_torch.py:575-587) -> Graph_test."""
    a1, a2 = np.load(path1), np.load(path2)
    if a1.shape != a2.shape:
        raise ValueError("the two layers must have the same number of nodes")
    return Graph_test(edges_from_dense(a1), edges_from_dense(a2), num_nodes=int(a1.shape[0]), cost=cost)


class GSet:
    """graph.GSet (U/graph.py:53-68)."""

    def __init__(self):
        self.graph_pool = {}

    def InsertGraph(self, gid, graph):
        assert gid not in self.graph_pool
        self.graph_pool[gid] = graph

    def Get(self, gid):
        assert gid in self.graph_pool
        return self.graph_pool[gid]

    def Sample(self):
        assert self.graph_pool
        gid = np.random.choice(list(self.graph_pool.keys()))
        return self.graph_pool[gid]

    def Clear(self):
        self.graph_pool.clear()


def build_csr(num_nodes, edges):
    """Symmetric CSR whose row order is the reference's in_edges insertion order
    (graph_struct.add_edge(x, y) then (y, x), U/PrepareBatchGraph.py:155-161): neighbours sorted by edge index."""
    e = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
    dst = np.stack([e[:, 1], e[:, 0]], 1).ravel()
    src = np.stack([e[:, 0], e[:, 1]], 1).ravel()
    order = np.argsort(dst, kind="stable")
    rowptr = np.zeros(num_nodes + 1, np.int64)
    np.cumsum(np.bincount(dst, minlength=num_nodes), out=rowptr[1:])
    return rowptr, src[order]


def _graph_csr(g, l):
    """build_csr of layer l of graph g, kept on the graph object (a replay batch draws the same few graphs over and over:
    the reference re-inserts every edge into a fresh graph_struct per batch, PrepareBatchGraph.py:142-161)."""
    cache = g.__dict__.setdefault("_mdb_csr", {})
    hit = cache.get(l)
    if hit is None or hit[0] is not g.edge_list[l]:
        hit = (g.edge_list[l],) + build_csr(g.num_nodes, g.edge_list[l])
        cache[l] = hit
    return hit[1], hit[2]


class GraphBatch:
    """B graphs resident in HBM as one block-diagonal multi-layer CSR (md_graphs_t).

    Global node id = g * stride + v with stride a multiple of 128, so embedding rows, labels and degrees are
    plain arrays indexed by global id and per-graph kernels never straddle graphs.
    """

    @staticmethod
    def prepare_host(graphs, cost=None):
        """Host side of the upload: block-diagonal CSR per layer (+ degree-cost node weights) as pinned tensors --
        the input data format of the device path."""
        graphs = list(graphs)
        cost = cost if cost is not None else graphs[0].cost
        B = len(graphs)
        L = graphs[0].num_layers
        nmax = max(g.num_nodes for g in graphs)
        # S >= nmax + 1: the alive-CSR row pointers of graph g live at aptr[g*S .. g*S + n_act] inside one
        # (B*S + 1)-entry array, so with n_act == S the end pointer of graph g would be the same word as the start
        # pointer of graph g + 1 (two CTAs writing different values).  One spare slot per graph rules that out.
        S = (nmax // TILE_M + 1) * TILE_M
        if B * S >= 2 ** 31:
            raise ValueError("batch too large for int32 node ids")
        rowptr, col, src, nnz = [], [], [], []
        for l in range(L):
            rp = np.zeros(B * S + 1, np.int64)
            cols = []
            pos = 0
            for gi, g in enumerate(graphs):
                r, c = _graph_csr(g, l)
                rp[gi * S:gi * S + g.num_nodes + 1] = r + pos
                rp[gi * S + g.num_nodes + 1:(gi + 1) * S + 1] = r[-1] + pos
                cols.append(c + gi * S)
                pos += int(r[-1])
            if pos >= 2 ** 31:
                raise ValueError("too many edges for int32 CSR offsets")
            nnz.append(pos)
            c = np.concatenate(cols) if pos else np.zeros(1, np.int64)
            rowptr.append(_pin(torch.from_numpy(rp.astype(np.int32))))
            col.append(_pin(torch.from_numpy(c.astype(np.int32))))
            sr = np.repeat(np.arange(B * S, dtype=np.int32), np.diff(rp)) if pos else np.zeros(1, np.int32)
            src.append(_pin(torch.from_numpy(sr)))
        gnnz = np.zeros(B, np.int64)
        for l in range(L):
            gnnz += 2 * np.array([g.num_edges[l] for g in graphs], np.int64)
        out = dict(B=B, L=L, S=S, rowptr=rowptr, col=col, src=src, nnz=nnz, cost=cost, max_graph_nnz=int(gnnz.max()), graphs=graphs,
                   n_nodes=_pin(torch.tensor([g.num_nodes for g in graphs], dtype=torch.int32)),
                   n_edges=_pin(torch.tensor([[g.num_edges[l] for g in graphs] for l in range(L)], dtype=torch.int32)))
        if cost == "degree":
            if L != 2:
                raise ValueError("degree-cost reward is defined for two layers (D/mvc_env.py:127-133)")
            cost_arr = np.zeros(B * S, np.float64)
            ws = [np.zeros(B * S, np.float32) for _ in range(L)]
            for gi, g in enumerate(graphs):
                if g.weights is None:
                    g.cal_degree()
                # D/mvc_env.py:129-130: Python sum over the weight dict (sequential f64 adds in node order)
                tot = [float(sum(g.weights[l].tolist())) for l in range(L)]
                cost_arr[gi * S:gi * S + g.num_nodes] = g.weights[0] / tot[0] + g.weights[1] / tot[1]
                for l in range(L):
                    ws[l][gi * S:gi * S + g.num_nodes] = g.weights[l].astype(np.float32)
            out["node_w"] = [_pin(torch.from_numpy(w)) for w in ws]
            out["node_cost"] = _pin(torch.from_numpy(cost_arr))
        return out

    def __init__(self, graphs, device=None, cost=None, host=None):
        if not torch.cuda.is_available():
            raise RuntimeError("mdcommunity_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib.lib()
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.graphs = list(graphs)
        B = len(self.graphs)
        assert B > 0
        L = self.graphs[0].num_layers
        assert all(g.num_layers == L for g in self.graphs)
        if L > _lib.MAX_LAYERS:
            raise ValueError("at most %d layers" % _lib.MAX_LAYERS)
        self.cost = cost if cost is not None else self.graphs[0].cost
        if host is None:
            host = self.prepare_host(self.graphs, self.cost)
        S = host["S"]
        assert host["B"] == B and host["L"] == L and host["cost"] == self.cost
        self.B, self.L, self.S, self.n_total = B, L, S, B * S
        dev = self.device
        self.nnz = list(host["nnz"])
        up = lambda t: t.to(dev, non_blocking=True)
        self.rowptr = [up(t) for t in host["rowptr"]]
        self.col = [up(t) for t in host["col"]]
        self.src = [up(t) for t in host["src"]]
        self.n_nodes, self.n_edges = up(host["n_nodes"]), up(host["n_edges"])
        self.max_rank = torch.ones(B, dtype=torch.int32, device=dev)
        self._upload_max_rank()
        self.node_w, self.node_cost = [], None
        if self.cost == "degree":
            self.node_w = [up(t) for t in host["node_w"]]
            self.node_cost = up(host["node_cost"])
        self.h2d_bytes = sum(t.numel() * t.element_size() for t in
                             host["rowptr"] + host["col"] + host["src"] + [host["n_nodes"], host["n_edges"]] +
                             (host["node_w"] + [host["node_cost"]] if self.cost == "degree" else []))
        st = _lib.MdGraphs()
        st.n_graphs, st.n_layers, st.stride, st.n_total = B, L, S, B * S
        st.max_graph_nnz = min(int(host["max_graph_nnz"]), 2 ** 31 - 1)
        for l in range(L):
            st.rowptr[l] = self.rowptr[l].data_ptr()
            st.col[l] = self.col[l].data_ptr()
            st.src[l] = self.src[l].data_ptr()
            if self.node_w:
                st.node_w[l] = self.node_w[l].data_ptr()
        st.n_nodes = self.n_nodes.data_ptr()
        st.n_edges = self.n_edges.data_ptr()
        st.max_rank = self.max_rank.data_ptr()
        st.node_cost = self.node_cost.data_ptr() if self.node_cost is not None else None
        self.c = st

    def _upload_max_rank(self):
        """The LMCC of the intact graphs, when already known (an earlier batch computed it; the reference computes it when it
        builds the Graph object, U/graph.py:30-37), is uploaded instead of being read back after s0 -- that read-back is a
        host synchronize per batch, which stalls every other batch in flight.  Unknown: BatchRollout.reset() fills it in."""
        known = [int(g.max_rank) for g in self.graphs]
        self._max_rank_set = all(r > 0 for r in known)
        if self._max_rank_set:
            self.max_rank.copy_(torch.tensor(known, dtype=torch.int32), non_blocking=True)

    def upload(self, host):
        """Refresh the resident CSR (and degree-cost node weights) from another prepared host batch of EXACTLY the same shape
        (B, stride, entries per layer): asynchronous copies from the pinned host tensors on the current stream, no
        allocation -- the device buffers, every BatchRollout bound to this batch and their captured CUDA graphs stay valid.
        The serving pattern: a pool of resident batches that new inputs are streamed into."""
        if (host["B"], host["L"], host["S"], host["cost"]) != (self.B, self.L, self.S, self.cost) or list(host["nnz"]) != self.nnz:
            raise ValueError("upload needs a host batch of the same shape (B, L, stride, cost, entries per layer)")
        with torch.cuda.device(self.device):
            for l in range(self.L):
                self.rowptr[l].copy_(host["rowptr"][l], non_blocking=True)
                self.col[l].copy_(host["col"][l], non_blocking=True)
                self.src[l].copy_(host["src"][l], non_blocking=True)
                if self.cost == "degree":
                    self.node_w[l].copy_(host["node_w"][l], non_blocking=True)
            self.n_nodes.copy_(host["n_nodes"], non_blocking=True)
            self.n_edges.copy_(host["n_edges"], non_blocking=True)
            if self.cost == "degree":
                self.node_cost.copy_(host["node_cost"], non_blocking=True)
            if int(host["max_graph_nnz"]) > int(self.c.max_graph_nnz) or \
                    max(g.num_nodes for g in host["graphs"]) > max(g.num_nodes for g in self.graphs):
                raise ValueError("a graph of the new batch is larger (nodes or CSR entries) than the resident batch was sized for")
            self.graphs = list(host["graphs"])
            self._upload_max_rank()
        return self

    def ptr(self):
        return C.byref(self.c)
