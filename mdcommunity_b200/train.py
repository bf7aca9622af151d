"""DQN training step, epsilon-greedy simulator and n-step replay memory on the device-resident environment
(SURVEY 8f N1 / N2).

Mirrors /root/reference/code/MultiDismantler_unit_cost:
  MultiDismantler_net.train_forward   MultiDismantler_net_graphsage.py:92-241
  MultiDismantler.fit / calc_loss     MultiDismantler_torch.py:381-431 (MSE + Alpha * Laplacian reconstruction, Adam)
  MultiDismantler.Fit                 :315-344 (n-step targets from the snapshot network)
  Run_simulator / PlayGame            :183-212
  NStepReplayMem                      nstep_replay_mem.py:18-110

What runs where.  The environment (node removal, LMCC, alive CSR, status counters), the batch preparation the reference
does in PrepareBatchGraph (compact re-indexing, COO of the alive edges, node / aux features) and the neighbour
aggregation (torch_sparse.spmm -> `ops.spmm`, forward and backward through md_spmm_csr_f32) are CUDA: the batch is
assembled from the env kernel's outputs with a handful of device-side index operations, no Python loop touches an edge.
The dense algebra of the 31 205-parameter network, the loss, their backward and the Adam update are hand-written too
(`dense.py` over csrc/md_train.cu: FP32 tile GEMM with a deterministic split over K for the weight gradients, fused
relu + L2-normalise, tanh, the two-layer fusion, the Q head and layer gate, forward AND backward; no cuBLAS, no torch
autograd formula).  Both layers and both row kinds (nodes, graph rows) are stacked into one [2(n+B), 64] matrix, so an
embedding iteration is 1 SpMM + 4 GEMMs + 1 normalise.  Data parallelism: every rank takes the round-robin shard of the replay batch, the losses are normalised by
the GLOBAL batch size / edge counts (one tiny all-reduce), so the SUM all-reduce of one flat gradient buffer gives
exactly the single-process gradient.

A replay entry does not copy Python sets (U/mvc_env.py:59-60 copies remove_edge every step): a state is a prefix of an
episode's removal sequence, and the pruned-edge sets are a function of the removed set (labels are recomputed by
md_env_set_state), so an entry is (episode, prefix length, action, n-step reward, successor prefix length, terminal).
"""
import random

import numpy as np
import torch
import torch.distributed as dist

from . import dense
from .graph import GraphBatch
from .rollout import BatchRollout, QWeights
from .sharding import shard_indices

GAMMA = 1            # U/MultiDismantler_torch.py:33
LEARNING_RATE = 1e-4  # :38
ALPHA = 0.001        # :40 weight of the reconstruction loss
N_STEP = 5           # :48
BATCH_SIZE = 64      # :54
MEMORY_SIZE = 100000  # :39
MAX_BP_ITER = 3      # :62

PARAM_SHAPES = {
    "w_n2l": (2, 64), "p_node_conv": (64, 64), "p_node_conv2": (64, 64), "p_node_conv3": (128, 64),
    "h1_weight": (64, 32), "h2_weight": (36, 1), "cross_product": (64, 1), "w_layer1": (64, 128), "w_layer2": (128, 1),
    "layerNodeAttention_weight.trans": (64, 64), "layerNodeAttention_weight.bias": (64,),
    "layerNodeAttention_weight.logis.parameter.weight": (1, 64), "layerNodeAttention_weight.logis.parameter.bias": (1,),
}


class QNet(torch.nn.Module):
    """The reference's parameter set under the reference's names (load_state_dict of its 14-key checkpoints works;
    `last_w` is the same tensor as `h2_weight`, net_graphsage.py:69).  Every Parameter is a view into ONE flat buffer
    (`flat`, offsets 16-byte aligned): the Adam kernel and the gradient all-reduce work on the flat buffers."""

    def __init__(self, device, seed=None):
        super().__init__()
        gen = torch.Generator().manual_seed(seed) if seed is not None else None
        self._names, self._slices, o = {}, {}, 0
        for k, shp in PARAM_SHAPES.items():
            k_el = int(np.prod(shp))
            self._slices[k] = (o, k_el)
            o += (k_el + 3) // 4 * 4
        self.flat = torch.zeros(o, dtype=torch.float32, device=device)
        self.flat_grad = torch.zeros(o, dtype=torch.float32, device=device)
        for k, shp in PARAM_SHAPES.items():
            if k.endswith("trans"):
                init = torch.eye(64)
            elif k.endswith(".bias") and "logis" not in k:
                init = torch.zeros(shp)
            else:  # net_graphsage.py:30: fmod(normal(0, std = 0.01), 2) -- effectively a plain N(0, 0.01): no wrap at 2
                init = torch.fmod(torch.normal(0.0, 0.01, size=tuple(shp), generator=gen), 2)
            name = k.replace(".", "__")
            self._names[k] = name
            lo, k_el = self._slices[k]
            view = self.flat[lo:lo + k_el].view(shp)
            view.copy_(init)
            self.register_parameter(name, torch.nn.Parameter(view))

    def p(self, key):
        return getattr(self, self._names[key])

    def params_in_key_order(self):
        return [self.p(k) for k in dense.PARAM_KEYS]

    def tensors(self):
        """reference key -> detached view of the parameter (what the dense kernels read)."""
        return {k: self.p(k).detach() for k in PARAM_SHAPES}

    def grad_tensors(self):
        """reference key -> view into flat_grad (what dense.backward writes)."""
        return {k: self.flat_grad[lo:lo + n_el].view(PARAM_SHAPES[k]) for k, (lo, n_el) in self._slices.items()}

    def state_dict_ref(self):
        sd = {k: self.p(k).detach().clone() for k in PARAM_SHAPES}
        sd["last_w"] = sd["h2_weight"]
        return sd

    def load_state_dict_ref(self, sd):
        with torch.no_grad():
            for k in PARAM_SHAPES:
                v = sd[k]
                v = torch.as_tensor(np.asarray(v)) if not isinstance(v, torch.Tensor) else v
                self.p(k).copy_(v.reshape(PARAM_SHAPES[k]).to(self.p(k).device))

    def rollout_weights(self):
        """The same parameters packed for the CUDA rollout kernels (Predict / greedy play)."""
        return QWeights({k: v.cpu() for k, v in self.state_dict_ref().items()}, self.p("w_n2l").device)


class TrainBatch:
    """PrepareBatchGraph.SetupTrain on the device: compact numbering of the active nodes of every graph, per-layer COO of
    the alive edges (row-sorted, neighbours in edge-list order), node features, aux features."""

    def __init__(self, ro: BatchRollout):
        b = ro.batch
        B, L, S, dev = b.B, b.L, b.S, b.device
        n_act = ro.n_act.long()
        off = torch.zeros(B + 1, dtype=torch.long, device=dev)
        off[1:] = torch.cumsum(n_act, 0)
        n = int(off[-1])
        self.B, self.L, self.n, self.off = B, L, n, off
        gid = torch.repeat_interleave(torch.arange(B, device=dev), n_act)
        pos = torch.arange(n, device=dev) - off[gid]
        slot = gid * S + pos
        v = ro.act_list[slot].long()                       # global node slot of every compact row
        self.gid, self.v = gid, v
        self.gidx = torch.full((b.n_total,), -1, dtype=torch.long, device=dev)
        self.gidx[v] = torch.arange(n, device=dev)
        self.index, self.node_input, self.edge_num = [], [], []
        cnts, cols = [], []
        for l in range(L):
            a0, a1 = ro.aptr[l][slot].long(), ro.aptr[l][slot + 1].long()
            cnt = a1 - a0
            row = torch.repeat_interleave(torch.arange(n, device=dev), cnt)
            start = torch.cumsum(cnt, 0) - cnt
            e = torch.arange(int(cnt.sum()), device=dev) - start[row] + a0[row]
            col = self.gidx[ro.acol[l][e].long()]
            self.index.append(torch.stack([row, col]))
            cnts.append(cnt), cols.append(col)
            self.edge_num.append(int(row.numel()))
            if b.cost == "degree":
                wv = b.node_w[l][v].float()
                x = torch.stack([wv, torch.ones_like(wv)], 1)
            else:  # degree / max degree of the graph (net_graphsage.py:100-112)
                d = ro.deg[l][v].float() / ro.max_deg[l][gid].float()
                x = torch.stack([d, d], 1)
            self.node_input.append(x)
        self.aux = ro.aux.clone()                          # [B, L, 4]
        self.N = n + B
        # operands of the hand-written network: block CSR (aggregation + sub-graph sum), its transpose, the Laplacian
        self.operands = dense.Operands.from_alive_csr(n, B, gid, off, n_act, cnts, cols, self.node_input, self.aux) if L == 2 else None

    def rows_of(self, local_ids, stride):
        """compact row of (graph i, local node id local_ids[i])."""
        ids = torch.as_tensor(np.asarray(local_ids, dtype=np.int64), device=self.gidx.device)
        rows = self.gidx[torch.arange(self.B, device=ids.device) * stride + ids]
        if bool((rows < 0).any()):
            raise ValueError("an action is not an active node of its state (PrepareBatchGraph.py:141)")
        return rows


def network(net: QNet, tb: TrainBatch):
    """q for every compact row [n, 1] and the fused, normalised embeddings [2, n + B, 64] (layer, node rows then graph
    rows) -- train_forward :92-241 and test_forward :243-394 share this arithmetic.  Runs dense.forward (hand-written
    kernels); differentiable through dense.backward."""
    assert tb.L == 2, "the reference network is hard-wired to two layers"
    return dense.network(net.params_in_key_order(), tb.operands)


def calc_loss(q_pred, fused, tb: TrainBatch, target, batch_global=None, edge_num_global=None, alpha=ALPHA):
    """MSE(target, q_pred) + alpha * sum_l 2 trace(H_l^T (D - A)_l H_l) / edge_num_l  (MultiDismantler_torch.py:410-431).
    batch_global / edge_num_global: normalisers of the whole (multi-rank) batch; default = this batch."""
    bg = float(batch_global if batch_global is not None else q_pred.shape[0])
    en = [float(e) for e in (edge_num_global if edge_num_global is not None else tb.edge_num)]
    return dense.loss(tb.operands, q_pred, fused, target.float().contiguous(), bg, en, alpha)


def set_states(ro: BatchRollout, covered):
    """Put graph i of the batch into the state 'nodes covered[i] removed' (Predict / SetupTrain hand-over,
    MultiDismantler_torch.py:220-234,263-302).  numCoveredEdges, which the reference recounts from covered / remove_edge
    in get_status_info (PrepareBatchGraph.py:45-52), is a function of the removal ORDER: edges to neighbours that were
    still in the same mutual component at removal time.  It is re-derived by replaying the prefix on the device."""
    b = ro.batch
    T = max([len(c) for c in covered] + [1])
    acts = np.full((b.B, T + 1), -1, np.int32)
    for i, c in enumerate(covered):
        acts[i, :len(c)] = np.asarray(list(c), dtype=np.int32)
    a = torch.from_numpy(acts).to(b.device)
    ro.reset()
    for t in range(T):
        ro.step(a, t)                                      # graphs whose prefix is shorter skip (action -1)
    return ro


class DQNTrainer:
    """fit() of the reference for one replay batch, data-parallel over torch.distributed when it is initialised."""

    def __init__(self, cost="unit", device=None, lr=LEARNING_RATE, seed=0):
        if not torch.cuda.is_available():
            raise RuntimeError("mdcommunity_b200.train needs a CUDA device (no CPU fallback)")
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.cost = cost
        self.net = QNet(self.device, seed)
        self.net_T = QNet(self.device, seed)
        self.TakeSnapShot()
        self.optimizer = dense.FlatAdam(self.net.flat, self.net.flat_grad, [self.net.p(k) for k in PARAM_SHAPES],
                                        self.net._slices, PARAM_SHAPES, lr)

    def LoadModel(self, state_dict):
        self.net.load_state_dict_ref(state_dict)
        self.TakeSnapShot()

    def TakeSnapShot(self):  # :312-313
        self.net_T.load_state_dict_ref(self.net.state_dict_ref())

    def _train_batch(self, g_list, covered):
        batch = GraphBatch(g_list, self.device, self.cost)
        ro = BatchRollout(batch, None, fused=False)
        set_states(ro, covered)
        return TrainBatch(ro), batch

    def train_forward(self, g_list, covered, actions):
        """(q_pred [B,1], cur_message_layer [2, n + B, 64], TrainBatch) for this process's graphs (autograd-enabled)."""
        tb, batch = self._train_batch(g_list, covered)
        q, fused = network(self.net, tb)
        rows = tb.rows_of(actions, batch.S)
        return dense.take_rows(q, rows), fused, tb

    def fit(self, g_list, covered, actions, list_target):
        """One optimizer step on the batch (MultiDismantler_torch.py:381-408).  Under torch.distributed every rank passes
        the SAME full batch and works on its round-robin shard; returns the global loss.  Forward, loss, backward and
        Adam are direct calls of the hand-written kernels (no autograd graph is built)."""
        world = dist.get_world_size() if dist.is_initialized() else 1
        rank = dist.get_rank() if dist.is_initialized() else 0
        idx = shard_indices(len(g_list), rank, world)
        tgt = torch.as_tensor(np.asarray(list_target, dtype=np.float64).reshape(-1, 1)[idx]).float().to(self.device)
        tb, batch = self._train_batch([g_list[i] for i in idx], [covered[i] for i in idx])
        o = tb.operands
        rows = tb.rows_of([actions[i] for i in idx], batch.S)
        en = torch.tensor(tb.edge_num, dtype=torch.float64, device=self.device)
        if world > 1:
            dist.all_reduce(en)
        params = self.net.tensors()
        c = dense.forward(params, o)
        q_pred = c["q"].index_select(0, rows)
        loss, lc = dense.loss_forward(o, q_pred, c["F"], tgt.reshape(-1), len(g_list), en.tolist(), ALPHA)
        dq_pred, dF = dense.loss_backward(o, lc)
        dq = torch.zeros(tb.n, dtype=torch.float32, device=self.device)
        dq.index_copy_(0, rows, dq_pred)
        self.optimizer.zero_grad()
        dense.backward(params, o, c, dq, dF, self.net.grad_tensors(), zeroed=True)
        self.optimizer.adopt_flat_grads()
        lv = loss.detach().double().reshape(1)
        if world > 1:
            dist.all_reduce(self.net.flat_grad)              # SUM of the shard gradients == single-process gradient
            dist.all_reduce(lv)
        self.optimizer.step()
        return float(lv[0])


def allreduce_gradients(params):
    """SUM all-reduce of ONE flat buffer holding every gradient (31 205 floats for the reference network)."""
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params])
    dist.all_reduce(flat)
    o = 0
    for p in params:
        k = p.numel()
        p.grad = flat[o:o + k].reshape(p.shape).clone()
        o += k
    return flat


# ---------------------------------------------------------------------------------------------------------------
# n-step replay memory and the epsilon-greedy simulator
# ---------------------------------------------------------------------------------------------------------------
class ReplaySample:
    def __init__(self):
        self.g_list, self.list_st, self.list_s_primes, self.list_at, self.list_rt, self.list_term = [], [], [], [], [], []


class NStepReplayMem:
    """nstep_replay_mem.py:18-110 with states stored as (episode actions, prefix length)."""

    def __init__(self, memory_size=MEMORY_SIZE):
        self.memory_size = memory_size
        self.items = [None] * memory_size
        self.current = 0
        self.count = 0

    def add(self, g, seq, t, a_t, r_t, t_prime, terminal):
        self.items[self.current] = (g, seq, t, a_t, r_t, t_prime, terminal)
        self.count = max(self.count, self.current + 1)
        self.current = (self.current + 1) % self.memory_size

    def add_episode(self, g, actions, rewards, n_step=N_STEP):
        """add_from_env (:57-80): actions / rewards of one finished episode."""
        T = len(actions)
        assert T > 0 and len(rewards) == T
        sums = [0.0] * T
        sums[T - 1] = rewards[T - 1]
        for i in range(T - 2, -1, -1):
            sums[i] = sums[i + 1] + rewards[i]
        seq = [int(a) for a in actions]
        for i in range(T):
            if i + n_step >= T:
                self.add(g, seq, i, seq[i], sums[i], T, True)
            else:
                self.add(g, seq, i, seq[i], sums[i] - sums[i + n_step], i + n_step, False)

    def sampling(self, batch_size, rng=random):
        assert self.count >= batch_size
        out = ReplaySample()
        for i in rng.sample(range(self.count), batch_size):
            g, seq, t, a, r, tp, term = self.items[i]
            out.g_list.append(g)
            out.list_st.append(seq[:t])
            out.list_s_primes.append(seq[:tp])
            out.list_at.append(a)
            out.list_rt.append(r)
            out.list_term.append(term)
        return out


class Simulator:
    """Run_simulator (:183-208) on the device environment: one episode per graph of the batch in lock step; per step ONE
    coin decides for all environments (as in the reference) between the network's arg-max and a uniformly random
    available node.  Random picks use the host RNG over the device's active-node lists."""

    def __init__(self, trainer: DQNTrainer, graphs, rng=None):
        self.trainer = trainer
        self.graphs = graphs
        self.batch = GraphBatch(graphs, trainer.device, trainer.cost)
        self.ro = BatchRollout(self.batch, trainer.net.rollout_weights(), fused=None)
        self.rng = rng if rng is not None else random.Random(0)

    def play(self, eps, mem: NStepReplayMem, n_step=N_STEP):
        """Plays every graph to termination, appends the transitions to `mem`, returns the number of episodes."""
        ro, b = self.ro, self.batch
        ro.w = self.trainer.net.rollout_weights()
        ro.reset()
        S = b.S
        while ro.num_done() < b.B:
            if self.rng.uniform(0, 1) >= eps:
                ro.forward()
                ro.step()
            else:
                n_act = ro.n_act.cpu().numpy()
                act = ro.act_list.cpu().numpy()
                done = ro.done.cpu().numpy()
                a = np.full((b.B, 1), -1, np.int32)
                for i in range(b.B):
                    if not done[i] and n_act[i] > 0:  # randomAction (U/mvc_env.py:88-100): uniform over the available nodes
                        a[i, 0] = act[i * S + self.rng.randrange(int(n_act[i]))] - i * S
                ro.step(torch.from_numpy(a).to(b.device), 0)
        res = ro.results()
        for i, g in enumerate(self.graphs):
            T = int(res["n_steps"][i])
            if T == 0:
                continue
            acts = res["actions"][i, :T]
            if b.cost == "unit":  # r_t = -LMCC / (LMCC_0 * N)  (U/mvc_env.py:133-137)
                rew = [-float(r) / (g.max_rank * float(g.num_nodes)) for r in res["ranks"][i, :T]]
            else:  # r_t = -LMCC / LMCC_0 * (w0[a]/W0 + w1[a]/W1) / 2  (D/mvc_env.py:127-133)
                tw = [float(sum(g.weights[l].tolist())) for l in range(2)]
                rew = [-float(r) / g.max_rank * (g.weights[0][a] / tw[0] + g.weights[1][a] / tw[1]) / 2.0
                       for a, r in zip(acts, res["ranks"][i, :T])]
            mem.add_episode(g, acts, rew, n_step)
        return len(self.graphs)


def fit_targets(trainer: DQNTrainer, sample: ReplaySample):
    """Fit (:315-344): target = r + GAMMA * max_a' Q_T(s', a') for non-terminal transitions."""
    B = len(sample.g_list)
    target = np.zeros((B, 1))
    live = [i for i in range(B) if not sample.list_term[i]]
    if live:
        batch = GraphBatch([sample.g_list[i] for i in live], trainer.device, trainer.cost)
        ro = BatchRollout(batch, None, fused=False)
        set_states(ro, [sample.list_s_primes[i] for i in live])
        with torch.no_grad():
            tb = TrainBatch(ro)
            q = dense.forward(trainer.net_T.tensors(), tb.operands)["q"]
            for j, i in enumerate(live):
                lo, hi = int(tb.off[j]), int(tb.off[j + 1])
                target[i, 0] = GAMMA * (float(q[lo:hi].max()) if hi > lo else 0.0)
    for i in range(B):
        target[i, 0] += sample.list_rt[i]
    return target
