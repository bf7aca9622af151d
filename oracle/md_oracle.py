"""CPU oracle for the MultiDismantler greedy-dismantling rollout.  TEST INFRASTRUCTURE ONLY.

This file is a CPU restatement (numpy / scipy / torch-CPU) of the reference's algorithm for the hot
path named in BASELINE.json.  It is imported only by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product (mdcommunity_b200/) never imports it.

Parity status: PINNED.  The restatement is checked against golden vectors produced by running the
unmodified reference in the build container (tests/golden/make_golden.py -> tests/golden/*.npz):
per-step LMCC, AUDC score, covered/removed edge counters, active sets, aux features (bit-exact) and
Q vectors (|dq| <= 2e-5*|q| + 1e-6), see tests/test_oracle_golden.py.

Citations are relative to /root/reference/code; U/ = MultiDismantler_unit_cost/, D/ = MultiDismantler_degree_cost/.

Third-party pieces restated here because they are not vendored in the reference tree:
  * torch_sparse.spmm (torch-sparse 0.6.18, uv.lock:1451-1500): out[row] += value * matrix[col]
    -> `spmm_coo` below (gather, multiply, index_add in COO order).
  * networkx.connected_components (networkx 3.x): only its result as a *partition* is used by
    U/Mcc.py:22-38, so scipy.sparse.csgraph.connected_components is an equivalent substitute.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.csgraph as csgraph
import torch
import torch.nn.functional as F

NEG_INF = 2147483647 / 2  # U/MultiDismantler_torch.py:60  (`inf`)
EMBED = 64
MAX_BP_ITER = 3  # U/MultiDismantler_torch.py:62


# ----------------------------------------------------------------------------------------------
# graph container  (U/graph.py:70-84 Graph_test, D/graph.py:39-51 cal_degree)
# ----------------------------------------------------------------------------------------------
class MultiplexGraph:
    """Static L-layer multiplex graph: per-layer undirected edge lists in the reference's G.edges() order."""

    def __init__(self, num_nodes, edge_lists, weights=None, max_rank=None):
        self.num_nodes = int(num_nodes)
        self.edges = [np.asarray(e, dtype=np.int64).reshape(-1, 2) for e in edge_lists]
        self.num_layers = len(self.edges)
        self.num_edges = [len(e) for e in self.edges]
        self.weights = None if weights is None else [np.asarray(w, dtype=np.float64) for w in weights]
        # U/graph.py:80-84 ori_rank: max_rank = LMCC of the intact graph
        self.max_rank = int(max_rank) if max_rank is not None else lmcc_size(
            mutual_labels(self, np.zeros(self.num_nodes, bool), np.zeros(self.num_nodes, np.int64)),
            np.zeros(self.num_nodes, bool))

    @staticmethod
    def degree_weights(num_nodes, edge_lists):
        """D/graph.py:39-51: w_l(v) = deg_l(v) / max_v deg_l(v), float64."""
        out = []
        for e in edge_lists:
            e = np.asarray(e, dtype=np.int64).reshape(-1, 2)
            deg = np.bincount(e.ravel(), minlength=num_nodes).astype(np.int64)
            out.append(deg / deg.max())
        return out


# ----------------------------------------------------------------------------------------------
# LMCC  (U/Mcc.py:3-38, called from U/mvc_env.py:140-162)
# ----------------------------------------------------------------------------------------------
def mutual_labels(g: MultiplexGraph, removed: np.ndarray, labels: np.ndarray) -> np.ndarray:
    """Mutually-connected-component labels (label = smallest node id of the block).

    Restates Mcc.MCC: alternate connected components per layer, deleting every edge that crosses
    a component of the other layer, until the partitions agree (U/Mcc.py:30-38).  The deleted edges
    (`remove_edge`) are exactly the edges between un-removed nodes with different final labels, so
    the label vector carries the reference's whole env state (SURVEY.md section 7 "State equivalence").
    `labels` is the previous partition (components only ever refine), all zeros for an intact graph.
    """
    n = g.num_nodes
    lab = labels.copy()
    stable = 0
    l = 0
    while stable < g.num_layers:
        e = g.edges[l]
        u, v = e[:, 0], e[:, 1]
        keep = (~removed[u]) & (~removed[v]) & (lab[u] == lab[v])
        a = sp.coo_matrix((np.ones(int(keep.sum()), np.int8), (u[keep], v[keep])), shape=(n, n))
        _, comp = csgraph.connected_components(a, directed=False)
        # canonical label: min node id in the component
        order = np.argsort(comp, kind="stable")
        first = np.full(comp.max() + 1, n, np.int64)
        np.minimum.at(first, comp, np.arange(n))
        new = first[comp]
        del order
        if np.array_equal(new, lab):
            stable += 1
        else:
            stable = 1
            lab = new
        l = (l + 1) % g.num_layers
    return lab


def lmcc_size(labels: np.ndarray, removed: np.ndarray) -> int:
    """Mcc.find_max_set_length (U/Mcc.py:12-13) over the components of the un-removed nodes."""
    alive = labels[~removed]
    if alive.size == 0:
        return 0
    return int(np.bincount(alive).max())


# ----------------------------------------------------------------------------------------------
# environment  (U/mvc_env.py:8-162, D/mvc_env.py:68-133)
# ----------------------------------------------------------------------------------------------
class OracleEnv:
    def __init__(self, variant="unit"):
        assert variant in ("unit", "degree")
        self.variant = variant
        self.graph = None

    def s0(self, g: MultiplexGraph):  # U/mvc_env.py:31-52
        self.graph = g
        n = g.num_nodes
        self.removed = np.zeros(n, bool)
        self.action_list = []
        self.numCoveredEdges = [0] * g.num_layers
        self.score = 0.0
        self.MaxCCList = [1]
        self.labels = mutual_labels(g, self.removed, np.zeros(n, np.int64))
        self.rank_list = []
        if self.variant == "degree":
            # D/mvc_env.py:129-130: Python sum over the weight dict in node order (sequential f64 adds)
            self._wsum = [float(sum(g.weights[l].tolist())) for l in range(g.num_layers)]

    def alive_edge_mask(self, l):
        e = self.graph.edges[l]
        u, v = e[:, 0], e[:, 1]
        return (~self.removed[u]) & (~self.removed[v]) & (self.labels[u] == self.labels[v])

    def alive_degree(self, l):
        e = self.graph.edges[l][self.alive_edge_mask(l)]
        return np.bincount(e.ravel(), minlength=self.graph.num_nodes)

    def num_removed_edges(self, l):
        """len(remove_edge[l]) / 2 of the reference.

        remove_edge only ever grows (U/Mcc.py:3-10): it holds every edge that was cut between two
        un-removed nodes of different blocks, including edges whose endpoint was removed later (those
        are skipped by the numCoveredEdges count, U/mvc_env.py:81-84).  Every edge is therefore in
        exactly one of {alive, covered, pruned}:  pruned = E - covered - alive.
        """
        return self.graph.num_edges[l] - self.numCoveredEdges[l] - int(self.alive_edge_mask(l).sum())

    def stepWithoutReward(self, a):  # U/mvc_env.py:74-87
        g = self.graph
        a = int(a)
        assert not self.removed[a]
        for l in range(g.num_layers):
            # neighbours still un-removed and not separated from `a` by an earlier prune (:81-84)
            self.numCoveredEdges[l] += int(self.alive_degree(l)[a])
        self.removed[a] = True
        self.action_list.append(a)
        self.labels = mutual_labels(g, self.removed, self.labels)
        rank = float(lmcc_size(self.labels, self.removed))
        self.rank_list.append(int(rank))
        if self.variant == "unit":
            # U/mvc_env.py:133-137 + :86-87
            r_t = -float(rank) / (g.max_rank * float(g.num_nodes))
            self.score += -1 * r_t
            self.MaxCCList.append(-1 * r_t * g.num_nodes)
        else:
            # D/mvc_env.py:127-133 + :79-81
            w0, w1 = g.weights[0], g.weights[1]
            r_t = -rank / (g.max_rank) * (float(w0[a]) / self._wsum[0] + float(w1[a]) / self._wsum[1]) / 2.0
            self.score += -1 * r_t
            self.MaxCCList.append(rank / (g.max_rank))
        return r_t

    step = stepWithoutReward

    def isTerminal(self):  # U/mvc_env.py:128-131  (E == covered + pruned  <=>  no alive edge in that layer)
        return any(int(self.alive_edge_mask(l).sum()) == 0 for l in range(self.graph.num_layers))

    # -- batch-prep quantities (U/PrepareBatchGraph.py:35-100) ---------------------------------
    def status(self):
        """active mask, per-layer alive degree, aux features [L,4] (f64 quotients cast to f32)."""
        g = self.graph
        degs = [self.alive_degree(l) for l in range(g.num_layers)]
        active = degs[0] > 0
        for l in range(1, g.num_layers):
            assert np.array_equal(active, degs[l] > 0)  # U/PrepareBatchGraph.py:73
        aux = np.zeros((g.num_layers, 4), np.float64)
        ncov = len(self.action_list)
        for l in range(g.num_layers):
            d = degs[l].astype(np.int64)
            twohop = int((d * (d - 1) // 2).sum())  # :62-72 == sum_v C(deg_alive(v), 2)
            aux[l, 0] = ncov / g.num_nodes
            aux[l, 1] = self.numCoveredEdges[l] / g.num_edges[l]
            aux[l, 2] = twohop / (g.num_nodes * g.num_nodes)
            aux[l, 3] = 1.0
        return active, degs, aux.astype(np.float32)


# ----------------------------------------------------------------------------------------------
# Q network  (U/MultiDismantler_net_graphsage.py:243-394, D/MultiDismantler_net.py:216-349,
#             U/MRGNN/mutil_layer_weight.py:266-313)
# ----------------------------------------------------------------------------------------------
def load_weights(npz_path):
    z = np.load(npz_path)
    return {k: torch.from_numpy(z[k].astype(np.float32)) for k in z.files}


def spmm_coo(row, col, m, matrix):
    """torch_sparse.spmm with all values 1.0 (aggregatorID = 0, U/PrepareBatchGraph.py:236-237)."""
    out = torch.zeros((m, matrix.shape[1]), dtype=matrix.dtype)
    return out.index_add_(0, row, matrix.index_select(0, col))


def _fuse(w, embs, p):
    """BitwiseMultipyLogis.forward for layer_predict = p (mutil_layer_weight.py:266-285, 294-301, 304-313).

    L-layer statement (metapath_number = L): T_k = tanh(E_k @ trans + bias); sem_k = T_k * T_p for every k (k = p gives
    T_p * T_p, :296-300); weight = softmax_k(sigmoid(Linear(sem_k))) over all L layers (:279-280); Z = sum_{k != p}
    weight_k T_k accumulated in ascending k from zeros (:281-284); X = T_p + Z (:285).  The reference allocates
    node_features_temp with a hard-coded leading 2 (:270), so it only RUNS for L = 2; for L = 2 this function performs
    the same torch ops in the same order as before the generalisation (pinned by the golden vectors)."""
    L = len(embs)
    t = [torch.tanh(torch.matmul(e, w["layerNodeAttention_weight.trans"]) + w["layerNodeAttention_weight.bias"])
         for e in embs]
    lw = w["layerNodeAttention_weight.logis.parameter.weight"]
    lb = w["layerNodeAttention_weight.logis.parameter.bias"]
    sem = torch.stack([t[k] * t[p] for k in range(L)], dim=1)  # [n, L, 64]
    a = torch.sigmoid(F.linear(sem, lw, lb)).squeeze(-1)  # [n, L]
    a = a.reshape(-1, L)
    alpha = F.softmax(a, dim=1)
    z = None
    for k in range(L):
        if k == p:
            continue
        term = alpha[:, k].unsqueeze(1) * t[k]
        z = term if z is None else z + term  # 0 + x == x exactly: the first add of the reference's zero-initialised Z
    return t[p] + z


F64_POOL_ROWS = 1 << 16  # graphs with more active rows than this: per-graph embedding sums in f64 (see forward_torch)
FUSE_CHUNK = 1 << 17  # rows per fusion call for very large graphs (bounds the [n, L, 64] temporaries; N = 1M only)


def forward_q(w, envs, variant, trace=None, head_rows=None):
    """test_forward over a batch of env states; returns (q [n_total] f32, active id list per graph).

    Unit cost builds node features from the alive degree (net_graphsage.py:248-261); degree cost
    uses [w_l(v), 1] (D/PrepareBatchGraph.py:133-136).  head_rows: evaluate fusion + Q head only on these compact row
    indices (q then has len(head_rows) entries) -- the N = 1M parity test samples rows this way.
    """
    with torch.no_grad():
        q, act_ids, off, _ = forward_torch(w, envs, variant, trace, head_rows=head_rows)
    return q[:, 0].numpy(), act_ids, off


def forward_torch(w, envs, variant, trace=None, head_rows=None):
    """The network as torch ops (differentiable in the weights): returns (q [n,1], active ids, offsets,
    dict(fused=[L x [n+B,64]], row=[L], col=[L], n=n)).  forward_q / train_loss are thin wrappers.

    L layers: the reference hard-wires lay_num = 2 (net_graphsage.py:251, :329-346, :393; mutil_layer_weight.py:270).
    Every parameter is layer-shared, so the L-layer network is defined as: the same per-layer
    embedding chain for each of the L layers, fusion softmax over L (see _fuse), Q mix q = sum_l softmax_l(gate) q_l
    accumulated in ascending l.  For L = 2 this is the reference's computation, op for op (golden-pinned); L = 3 is an
    extension with no reference output to compare with ("parity unpinned" for L = 3: the CUDA path is checked against
    this restatement only)."""
    L = envs[0].graph.num_layers
    assert L >= 2
    B = len(envs)
    stats = [e.status() for e in envs]
    act_ids = [np.nonzero(s[0])[0] for s in stats]
    cnt = [len(a) for a in act_ids]
    off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    n = int(off[-1])
    aux = torch.from_numpy(np.stack([s[2] for s in stats]))  # [B, L, 4]
    gid = torch.from_numpy(np.repeat(np.arange(B), cnt))
    node_emb, y_emb = [], []
    coo_rows, coo_cols = [], []
    for l in range(L):
        rows, cols, x = [], [], []
        for b, e in enumerate(envs):
            g = e.graph
            idx = np.full(g.num_nodes, -1, np.int64)
            idx[act_ids[b]] = np.arange(cnt[b]) + off[b]  # ascending original id (PrepareBatchGraph.py:128-138)
            ed = g.edges[l][e.alive_edge_mask(l)]
            xs, ys = idx[ed[:, 0]], idx[ed[:, 1]]
            # graph_struct.add_edge(x,y) then (y,x); in_edges.head[i] keeps insertion order (:155-161, 217-249)
            dst = np.stack([ys, xs], 1).ravel()
            src = np.stack([xs, ys], 1).ravel()
            order = np.argsort(dst, kind="stable")
            rows.append(dst[order])
            cols.append(src[order])
            if variant == "unit":
                d = torch.from_numpy(stats[b][1][l][act_ids[b]].astype(np.float32))
                if d.numel():
                    x.append(torch.stack([d / d.max(), d / d.max()], 1))
            else:
                wv = torch.tensor(g.weights[l][act_ids[b]].tolist()).type(torch.FloatTensor)
                x.append(torch.stack([wv, torch.ones_like(wv)], 1))
        row = torch.from_numpy(np.concatenate(rows)) if rows else torch.zeros(0, dtype=torch.long)
        col = torch.from_numpy(np.concatenate(cols)) if cols else torch.zeros(0, dtype=torch.long)
        node_input = torch.cat(x, 0) if x else torch.zeros((0, 2))
        coo_rows.append(row)
        coo_cols.append(col)
        h = F.normalize(torch.relu(torch.matmul(node_input, w["w_n2l"])), p=2, dim=1)
        y = F.normalize(torch.relu(torch.matmul(torch.ones((B, 2)), w["w_n2l"])), p=2, dim=1)
        if trace is not None:
            trace.setdefault("h", {})[(l, 0)] = h.detach().numpy().copy()
            trace.setdefault("y", {})[(l, 0)] = y.detach().numpy().copy()
        for _ in range(MAX_BP_ITER):
            pool = spmm_coo(row, col, n, h)
            node_linear = torch.matmul(pool, w["p_node_conv"])
            if n > F64_POOL_ROWS:
                # torch_sparse.spmm(subgsum) is a sequential f32 scatter-add over the graph's rows.  Over ~1M rows (many of
                # them bit-identical: the first embedding depends on the degree only) its rounding errors no longer
                # average out -- the f32 running sum is quantised to ulp(1e5) ~ 0.008 per add and the result depends on
                # the row order at the 1e-2 level.  The reference cannot run such graphs at all (dense [n, n] adjacency);
                # for them the oracle forms the segment sum in f64, the exact value every f32 summation order approximates.
                y_pool = torch.zeros((B, h.shape[1]), dtype=torch.float64).index_add_(0, gid, h.double()).float()
            else:
                y_pool = spmm_coo(gid, torch.arange(n), B, h)
            y_linear = torch.matmul(y_pool, w["p_node_conv"])
            merged = torch.cat([node_linear, torch.matmul(h, w["p_node_conv2"])], 1)
            h = F.normalize(torch.relu(torch.matmul(merged, w["p_node_conv3"])), p=2, dim=1)
            y_merged = torch.cat([y_linear, torch.matmul(y, w["p_node_conv2"])], 1)
            y = F.normalize(torch.relu(torch.matmul(y_merged, w["p_node_conv3"])), p=2, dim=1)
            if trace is not None:
                trace["h"][(l, _ + 1)] = h.detach().numpy().copy()
                trace["y"][(l, _ + 1)] = y.detach().numpy().copy()
                trace.setdefault("ypool", {})[(l, _)] = y_pool.detach().numpy().copy()
        node_emb.append(h)
        y_emb.append(y)
    embs = [torch.cat([node_emb[l], y_emb[l]], 0) for l in range(L)]
    sel = None
    if head_rows is not None:  # fusion and Q head are row-wise: restrict them to the sampled node rows + all graph rows
        sel = torch.as_tensor(np.asarray(head_rows, dtype=np.int64))
        keep = torch.cat([sel, torch.arange(n, n + B)])
        embs = [e.index_select(0, keep) for e in embs]
        gid = gid.index_select(0, sel)
        n_rows = int(sel.numel())
    else:
        n_rows = n
    if n_rows + B > FUSE_CHUNK:
        fused = [torch.cat([F.normalize(_fuse(w, [e[c0:c0 + FUSE_CHUNK] for e in embs], p), p=2, dim=1)
                            for c0 in range(0, n_rows + B, FUSE_CHUNK)], 0) for p in range(L)]
    else:
        fused = [F.normalize(_fuse(w, embs, p), p=2, dim=1) for p in range(L)]
    if trace is not None:
        trace["fused"] = [f.detach().numpy().copy() for f in fused]
        trace["n"] = n
    qs, gates = [], []
    n_full, n = n, n_rows
    for l in range(L):
        hf, yf = fused[l][:n], fused[l][n:]
        rep_y = yf.index_select(0, gid)
        # literal op order of :356-363: outer product [n,64,64] first, then @ cross_product (chunked)
        e_sa = torch.empty_like(hf)
        for c0 in range(0, n, 4096):
            t1 = torch.matmul(hf[c0:c0 + 4096].unsqueeze(2), rep_y[c0:c0 + 4096].unsqueeze(1))
            e_sa[c0:c0 + 4096] = torch.matmul(t1, w["cross_product"]).reshape(-1, EMBED)
        hidden = torch.relu(torch.matmul(e_sa, w["h1_weight"]))
        rep_aux = aux[:, l, :].index_select(0, gid)
        qs.append(torch.matmul(torch.cat([hidden, rep_aux], 1), w["h2_weight"]))
        gates.append(torch.matmul(torch.relu(torch.matmul(rep_y, w["w_layer1"])), w["w_layer2"]))
    sm = F.softmax(torch.cat(gates, 1), dim=1)
    q = sm[:, 0:1] * qs[0] + sm[:, 1:2] * qs[1]  # net_graphsage.py:393
    for l in range(2, L):
        q = q + sm[:, l:l + 1] * qs[l]
    if trace is not None:
        trace["gate_softmax"] = sm.detach().numpy().copy()
        trace["q_layers"] = [x.detach().numpy().copy() for x in qs]
    return q, act_ids, off, dict(fused=fused, row=coo_rows, col=coo_cols, n=n_full)


def predict(w, envs, variant):
    """MultiDismantler.Predict (U/MultiDismantler_torch.py:263-302): scatter q to [N] f64, -inf elsewhere."""
    q, act_ids, off = forward_q(w, envs, variant)
    out = []
    for b, e in enumerate(envs):
        p = np.full(e.graph.num_nodes, -NEG_INF, np.float64)
        p[act_ids[b]] = q[off[b]:off[b + 1]]
        p[np.asarray(e.action_list, dtype=np.int64)] = -NEG_INF
        out.append(p)
    return out


def arg_first_max(p):
    """argMax of U/MultiDismantler_torch.py:799-807: first index of the maximum.

    GetSol uses np.argsort(-pred)[:step] (:769) whose tie order is an artefact of numpy's introsort;
    ties are resolved here (and in the CUDA path) to the lowest node id.
    """
    return int(np.argmax(p))


def rollout(w, g, variant, step=1, forced_actions=None, record=False):
    """GetSol / GetSolution (U/MultiDismantler_torch.py:759-784, 711-736) for one graph.

    forced_actions: teacher forcing -- use the given action at step t instead of the arg-max.
    Returns dict(score, sol, ranks, maxcc, [q, active per step]).
    """
    env = OracleEnv(variant)
    env.s0(g)
    sol, qs, acts = [], [], []
    t = 0
    while not env.isTerminal():
        p = predict(w, [env], variant)[0]
        if record:
            active, _, _ = env.status()
            ids = np.nonzero(active)[0]
            acts.append(ids)
            qs.append(p[ids].astype(np.float32))
        order = np.lexsort((np.arange(len(p)), -p))[:step]  # q descending, node id ascending
        for a in order:
            if env.isTerminal():
                break
            if forced_actions is not None:
                a = forced_actions[t]
            env.stepWithoutReward(int(a))
            sol.append(int(a))
            t += 1
    return dict(score=env.score, sol=sol, ranks=list(env.rank_list), maxcc=list(env.MaxCCList),
                q=qs, active=acts, env=env)


# ----------------------------------------------------------------------------------------------
# heuristic baselines on the same environment
#   HDA : U/baseline/HDA/hda_2max.py:50-69 (delnode: key = max(deg_1, deg_2) in the pruned layers),
#         hda_add.py:50-69 (key = deg_1 + deg_2)
#   CI  : U/baseline/CI/ci_max.py:52-90 (basic_ci: (deg(v) - 1) * sum_{u in N(v)} (deg(u) - 1), -1 for
#         isolated nodes; key = max(ci_1, ci_2)), ci_add.py (key = ci_1 + ci_2)
#   loop: critical_number (hda_2max.py:71-93): remove the arg-max, MCC, until LMCC == 1;
#         score += LMCC / (LMCC_0 * N)  (same as the unit-cost AUDC, U/mvc_env.py:133-137)
# Tie rule: `sorted(..., reverse=True)[0]` is stable, i.e. the FIRST node in G.nodes() order (= lowest id) among
# the maxima.  The *_add.py scripts break ties with random.choice instead; the restatement (and the CUDA path)
# uses the lowest id there too, which is one admissible outcome of that random choice.
# ----------------------------------------------------------------------------------------------
def heuristic_scores(env: OracleEnv, kind: str, combine: str) -> np.ndarray:
    """Selection key of every node (int64), -1 for removed nodes / nodes the key cannot select."""
    g = env.graph
    per_layer = []
    for l in range(g.num_layers):
        d = env.alive_degree(l).astype(np.int64)
        if kind == "hda":
            s = d.copy()
        else:
            e = g.edges[l][env.alive_edge_mask(l)]
            nb = np.zeros(g.num_nodes, np.int64)  # sum over alive neighbours of (deg - 1)
            np.add.at(nb, e[:, 0], d[e[:, 1]] - 1)
            np.add.at(nb, e[:, 1], d[e[:, 0]] - 1)
            s = np.where(d == 0, -1, nb * (d - 1))
        per_layer.append(s)
    key = per_layer[0].copy()
    for s in per_layer[1:]:
        key = np.maximum(key, s) if combine == "max" else key + s
    key[env.removed] = np.iinfo(np.int64).min
    return key


def heuristic_rollout(g: MultiplexGraph, kind="hda", combine="max"):
    """critical_number of the baseline scripts on the oracle environment.  Returns dict(score, sol, ranks, maxcc)."""
    assert kind in ("hda", "ci") and combine in ("max", "add")
    env = OracleEnv("unit")
    env.s0(g)
    sol = []
    while not env.isTerminal():
        key = heuristic_scores(env, kind, combine)
        a = int(np.argmax(key))  # first (lowest id) maximum
        env.stepWithoutReward(a)
        sol.append(a)
    return dict(score=env.score, sol=sol, ranks=list(env.rank_list), maxcc=list(env.MaxCCList), env=env)


# ----------------------------------------------------------------------------------------------
# DQN training step
#   fit        U/MultiDismantler_torch.py:381-408  (SetupTrain, train_forward, calc_loss, backward, Adam step)
#   train_forward U/MultiDismantler_net_graphsage.py:92-241: the same network as test_forward; Q is read only at
#              the chosen action's row of each graph (action_select spmm, :199-203)
#   calc_loss  :410-431: MSE(target, q_pred) + Alpha * sum_l 2 trace(H_l^T Lap_l H_l) / edge_num_l with
#              Lap_l = D - A of the alive layer (PrepareBatchGraph.n2n_construct :217-249), H_l = the fused, normalised
#              node embeddings returned as cur_message_layer, edge_num_l = sum of the n2n values (directed alive entries)
# ----------------------------------------------------------------------------------------------
ALPHA = 0.001  # U/MultiDismantler_torch.py:40
LEARNING_RATE = 0.0001  # :38
GAMMA = 1  # :33
N_STEP = 5  # :48


def train_loss(w, envs, actions, targets, variant="unit"):
    """loss, q_pred [B,1], [H_0, H_1] for a batch of (state, action, target); w may hold requires_grad tensors."""
    q, act_ids, off, ex = forward_torch(w, envs, variant)
    rows = []
    for b, (ids, a) in enumerate(zip(act_ids, actions)):
        pos = np.nonzero(ids == int(a))[0]
        assert pos.size == 1, "the action must be an active node of its state (PrepareBatchGraph.py:141)"
        rows.append(int(off[b] + pos[0]))
    q_pred = q.index_select(0, torch.tensor(rows, dtype=torch.long))
    n = ex["n"]
    loss2 = torch.zeros(1)
    embs = []
    for l in range(len(ex["fused"])):
        h = ex["fused"][l][:n]
        embs.append(h)
        row, col = ex["row"][l], ex["col"][l]
        deg = torch.bincount(row, minlength=n).to(h.dtype)
        lap_h = deg.unsqueeze(1) * h - spmm_coo(row, col, n, h)
        recons = 2 * torch.trace(torch.matmul(h.t(), lap_h))
        loss2 = loss2 + recons / float(row.numel())
    tgt = torch.as_tensor(np.asarray(targets, dtype=np.float64).reshape(-1, 1)).type(torch.FloatTensor)
    loss1 = F.mse_loss(tgt, q_pred)
    return loss1 + ALPHA * loss2, q_pred, embs


def nstep_returns(rewards, n_step=N_STEP):
    """NStepReplayMem.add_from_env (U/nstep_replay_mem.py:57-80): per step i the n-step reward sum, the index of the
    successor state (None = terminal) -- states are prefixes of the removal sequence."""
    T = len(rewards)
    sums = [0.0] * T
    sums[T - 1] = rewards[T - 1]
    for i in range(T - 2, -1, -1):
        sums[i] = sums[i + 1] + rewards[i]
    out = []
    for i in range(T):
        if i + n_step >= T:
            out.append((sums[i], None, True))
        else:
            out.append((sums[i] - sums[i + n_step], i + n_step, False))
    return out
