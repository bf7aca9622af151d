"""Device-resident lock-step greedy rollout over a batch of graphs.

This is the B200 replacement for the reference's hot loop (GetSol / GetSolution / Test,
/root/reference/code/MultiDismantler_unit_cost/MultiDismantler_torch.py:711-784): every graph of the batch advances
one removal per iteration; graphs never interact (segment sums, max degree, arg-max and LMCC are all per graph),
so batching B independent rollouts is exactly B reference rollouts.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .graph import GraphBatch


def _on_device(fn):
    """Run a BatchRollout method with the batch's device current: the C ABI launches on the CURRENT device, and the
    stream handed over belongs to batch.device."""
    import functools

    @functools.wraps(fn)
    def wrapped(self, *a, **kw):
        with torch.cuda.device(self.batch.device):
            return fn(self, *a, **kw)
    return wrapped


class QWeights:
    """The reference's 14-tensor state_dict on the device, plus the two merged products the fused kernel uses."""

    KEYS = ["w_n2l", "p_node_conv", "p_node_conv2", "p_node_conv3", "h1_weight", "h2_weight", "cross_product",
            "w_layer1", "w_layer2", "layerNodeAttention_weight.trans", "layerNodeAttention_weight.bias",
            "layerNodeAttention_weight.logis.parameter.weight", "layerNodeAttention_weight.logis.parameter.bias"]

    def __init__(self, state_dict, device):
        sd = {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in state_dict.items()}
        for k in self.KEYS:
            if k not in sd:
                raise KeyError("checkpoint lacks %s (older 10-key checkpoints without fusion/gate weights are not usable)" % k)
        f = lambda a: np.ascontiguousarray(a, dtype=np.float32)
        c1, c2, c3 = sd["p_node_conv"].astype(np.float64), sd["p_node_conv2"].astype(np.float64), sd["p_node_conv3"].astype(np.float64)
        host = {
            "w_n2l": f(sd["w_n2l"]), "conv1": f(sd["p_node_conv"]), "conv2": f(sd["p_node_conv2"]), "conv3": f(sd["p_node_conv3"]),
            # [P@c1, H@c2] @ c3 == P @ (c1 @ c3[:64]) + H @ (c2 @ c3[64:]); products formed in f64, rounded once
            "w13": f(c1 @ c3[:64]), "w23": f(c2 @ c3[64:]),
            "trans": f(sd["layerNodeAttention_weight.trans"]), "bias": f(sd["layerNodeAttention_weight.bias"]),
            "logis_w": f(sd["layerNodeAttention_weight.logis.parameter.weight"]).reshape(-1),
            "logis_b": f(sd["layerNodeAttention_weight.logis.parameter.bias"]).reshape(-1),
            "cross": f(sd["cross_product"]).reshape(-1), "h1": f(sd["h1_weight"]), "h2": f(sd["h2_weight"]).reshape(-1),
            "wl1": f(sd["w_layer1"]), "wl2": f(sd["w_layer2"]).reshape(-1),
        }
        # one flat 16-byte aligned buffer
        offs, total = {}, 0
        for k, a in host.items():
            offs[k] = total
            total += (a.size + 3) // 4 * 4
        flat = np.zeros(total, np.float32)
        for k, a in host.items():
            flat[offs[k]:offs[k] + a.size] = a.ravel()
        self.flat = torch.from_numpy(flat).to(device)
        self.c = _lib.MdWeights()
        for k in host:
            setattr(self.c, k, self.flat.data_ptr() + 4 * offs[k])

    def ptr(self):
        return C.byref(self.c)


class BatchRollout:
    """Owns the mutable env state + activations for one GraphBatch and drives the kernels.

    results after run(): score [B] f64 (AUDC), n_steps [B], actions [B, T] (-1 padded), ranks [B, T], maxcc [B, T].
    """

    def __init__(self, batch: GraphBatch, weights: QWeights = None, max_steps=None, fused=None, tail=None):
        """fused: None = use md_embed_fused whenever the batch qualifies (stride <= md_fused_max_stride());
        False = always the unfused md_embed_init + md_embed_iter kernels (A/B tests, large graphs).
        tail: True = run() hands graphs that have shrunk to <= md_tail_max_nodes() active nodes over to md_rollout_tail
        (one CTA finishes each of them on chip); None / False = pure lock step to the end (the default: measured on C2 the
        tail kernel shortens a single batch by < 1 % and gains nothing once several batches are in flight)."""
        self.batch = batch
        self.lib = batch.lib
        self.w = weights
        dev = batch.device
        B, L, S, nt = batch.B, batch.L, batch.S, batch.n_total
        self.max_steps = int(max_steps if max_steps is not None else max(g.num_nodes for g in batch.graphs))
        i32 = lambda *s: torch.zeros(*s, dtype=torch.int32, device=dev)
        self.label = i32(nt)
        self.deg = [i32(nt) for _ in range(L)]
        self.act_list = i32(nt)
        self.aptr = [i32(nt + 1) for _ in range(L)]
        self.acol = [i32(max(1, batch.nnz[l])) for l in range(L)]
        self.asrc = [i32(max(1, batch.nnz[l])) for l in range(L)]
        self.n_act, self.n_removed, self.rank, self.done, self.n_steps = i32(B), i32(B), i32(B), i32(B), i32(B)
        self.cov_edges, self.alive_edges, self.max_deg = i32(L, B), i32(L, B), i32(L, B)
        self.twohop = torch.zeros(L, B, dtype=torch.int64, device=dev)
        self.score = torch.zeros(B, dtype=torch.float64, device=dev)
        self.aux = torch.zeros(B, L, 4, dtype=torch.float32, device=dev)
        self.best = torch.zeros(B, dtype=torch.int64, device=dev)
        self.trace_action = torch.full((B, self.max_steps), -1, dtype=torch.int32, device=dev)
        self.trace_rank = i32(B, self.max_steps)
        self.trace_maxcc = torch.zeros(B, self.max_steps, dtype=torch.float64, device=dev)
        self.n_done = i32(1)
        self.large = not self.lib.md_env_uses_small_path(batch.ptr())
        self.parent = i32(nt) if self.large else None
        self.comp_size = i32(3 * nt) if self.large else None
        self.lg_scratch = i32(9 * nt) if self.large else None  # incremental LMCC update of the multi-CTA env kernel
        self.flags = i32(32)
        self.cur_action = i32(B)
        self.tail_n = i32(B)
        s = _lib.MdState()
        s.label = self.label.data_ptr()
        for l in range(L):
            s.deg[l] = self.deg[l].data_ptr()
            s.aptr[l] = self.aptr[l].data_ptr()
            s.acol[l] = self.acol[l].data_ptr()
            s.asrc[l] = self.asrc[l].data_ptr()
        for name in ("act_list", "n_act", "n_removed", "cov_edges", "alive_edges", "max_deg", "twohop", "rank", "done",
                     "n_steps", "score", "aux", "best", "trace_action", "trace_rank", "trace_maxcc", "n_done", "flags",
                     "cur_action"):
            setattr(s, name, getattr(self, name).data_ptr())
        s.parent = self.parent.data_ptr() if self.large else None
        s.comp_size = self.comp_size.data_ptr() if self.large else None
        s.lg_scratch = self.lg_scratch.data_ptr() if self.large else None
        s.max_steps = self.max_steps
        s.cost_mode = 1 if batch.cost == "degree" else 0
        s.tail_n = None
        self.s = s
        # the same state with the hand-over to md_rollout_tail switched on: only run() uses it (teacher-forced steps,
        # forward() / q_values() inspection, top-k and heuristic rollouts keep every graph in the lock-step kernels)
        self.tail = bool(tail) and bool(self.lib.md_tail_applies(batch.ptr()))
        if tail and not self.tail:
            raise ValueError("md_rollout_tail handles two-layer batches on the single-CTA env path only")
        self.s_tail = _lib.MdState.from_buffer_copy(s)
        self.s_tail.tail_n = self.tail_n.data_ptr()
        # activations
        self.h = [[torch.zeros(nt, _lib.EMBED, dtype=torch.float32, device=dev) for _ in range(L)] for _ in range(2)]
        self.psum = torch.zeros(int(self.lib.md_psum_floats(L, nt)), dtype=torch.float32, device=dev)
        self.gconst = torch.zeros(B, _lib.GCONST, dtype=torch.float32, device=dev)
        self.q = torch.zeros(nt, dtype=torch.float32, device=dev)
        k = _lib.MdWork()
        for i in range(2):
            for l in range(L):
                k.h[i][l] = self.h[i][l].data_ptr()
        k.psum, k.gconst, k.q = self.psum.data_ptr(), self.gconst.data_ptr(), self.q.data_ptr()
        self.fused = (fused is not False) and bool(self.lib.md_fused_applies(batch.ptr()))
        if fused and not self.fused:
            raise ValueError("md_embed_fused handles strides up to %d (this batch: %d) and < 65536 CSR entries per graph" % (self.lib.md_fused_max_stride(), S))
        if self.fused:
            nbytes = int(self.lib.md_fused_ws_bytes(batch.ptr()))
            self.fused_ws = torch.zeros(nbytes + 1024, dtype=torch.uint8, device=dev)
            k.fused_ws = (self.fused_ws.data_ptr() + 1023) // 1024 * 1024
        else:
            k.fused_ws = None
        self.k = k
        self._pinned_done = torch.zeros(1, dtype=torch.int32).pin_memory()
        self._graph, self._graph_chunk, self._graph_w = None, 0, None
        self.use_graph = True
        self._runs = 0
        self.kernel_launches = 0
        self.reset()

    # -- helpers ----------------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.batch.device).cuda_stream)

    def _g(self):
        return self.batch.ptr()

    @_on_device
    def reset(self):
        """MvcEnv.s0 for every graph; also fills graph.max_rank (Graph_test.ori_rank) on first use."""
        self.trace_action.fill_(-1)
        self.tail_n.zero_()
        _lib.check(self.lib.md_env_reset(self._g(), C.byref(self.s), self._stream()), "md_env_reset")
        self.kernel_launches += 1
        if not getattr(self.batch, "_max_rank_set", False):
            self.batch.max_rank.copy_(self.rank)
            mr = self.rank.cpu().numpy()
            for g, r in zip(self.batch.graphs, mr):
                g.max_rank = int(r)
            self.batch._max_rank_set = True

    @_on_device
    def set_state(self, removed_mask):
        """Arbitrary state: removed_mask uint8 [n_total] on the device; self.cov_edges must hold numCoveredEdges."""
        assert removed_mask.dtype == torch.uint8 and removed_mask.is_cuda and removed_mask.numel() == self.batch.n_total
        _lib.check(self.lib.md_env_set_state(self._g(), C.byref(self.s), C.c_void_p(removed_mask.data_ptr()),
                                             self._stream()), "md_env_set_state")
        self.kernel_launches += 1

    @_on_device
    def forward(self):
        """One Q-network evaluation for all non-terminal graphs (test_forward + the Predict mask + arg-max key)."""
        lib, g, s, w, k, st = self.lib, self._g(), C.byref(self.s), self.w.ptr(), C.byref(self.k), self._stream()
        if self.fused:
            _lib.check(lib.md_embed_fused(g, s, w, k, st), "md_embed_fused")
            self.kernel_launches += 2
        else:
            _lib.check(lib.md_embed_init(g, s, w, k, st), "md_embed_init")
            for it in range(_lib.BP_ITERS):
                _lib.check(lib.md_embed_iter(g, s, w, k, it, st), "md_embed_iter")
            self.kernel_launches += 1 + _lib.BP_ITERS
        _lib.check(lib.md_graph_embed(g, s, w, k, st), "md_graph_embed")
        _lib.check(lib.md_qhead(g, s, w, k, st), "md_qhead")
        self.kernel_launches += 2

    @_on_device
    def step(self, actions=None, index=-1):
        """stepWithoutReward for every non-terminal graph.  actions: None -> arg-max of the last forward();
        int32 device tensor [B, stride] of local node ids -> column `index` (or column n_steps[g] when index < 0)."""
        if actions is None:
            rc = self.lib.md_env_step(self._g(), C.byref(self.s), None, 0, -1, self._stream())
        else:
            assert actions.dtype == torch.int32 and actions.is_cuda and actions.is_contiguous()
            rc = self.lib.md_env_step(self._g(), C.byref(self.s), C.c_void_p(actions.data_ptr()), actions.shape[1],
                                      index, self._stream())
        _lib.check(rc, "md_env_step")
        self.kernel_launches += 1

    def num_done(self):
        self._pinned_done.copy_(self.n_done, non_blocking=True)
        torch.cuda.current_stream(self.batch.device).synchronize()
        return int(self._pinned_done[0])

    def num_left(self):
        """Graphs that have left the lock-step iteration (terminal, or handed over to md_rollout_tail)."""
        self._pinned_done.copy_(self.flags[8:9], non_blocking=True)
        torch.cuda.current_stream(self.batch.device).synchronize()
        return int(self._pinned_done[0])

    # -- greedy rollout: pieces shared by run() and MultiStreamRollout -------------------------------------------
    def _run_state(self, step=1):
        return self.s_tail if (self.tail and step == 1) else self.s

    def _begin_run(self, step=1):
        """Start of a greedy rollout from the current env state: graphs that are already small go to md_rollout_tail."""
        if self.tail and step == 1:
            st = self._stream()
            _lib.check(self.lib.md_tail_handover(self._g(), C.byref(self.s_tail), st), "md_tail_handover")
            _lib.check(self.lib.md_rollout_tail(self._g(), C.byref(self.s_tail), self.w.ptr(), st), "md_rollout_tail")
            self.kernel_launches += 2

    def _enqueue_chunk(self, chunk, first):
        """`chunk` lock-step iterations on the current stream (one CUDA-graph launch from the second run of this object
        on), then -- tail mode -- one md_rollout_tail launch for the graphs the chunk handed over."""
        s = self._run_state()
        if self._graph is None and self._runs > 0 and not first and not self.large and self.use_graph:
            # a chunk of iterations as ONE CUDA graph, built when the rollout object is re-used (instantiation costs
            # about as much as one small rollout); the first chunk ran eagerly: one-time kernel attributes are set
            ex = C.c_void_p()
            _lib.check(self.lib.md_rollout_graph_create(self._g(), C.byref(s), self.w.ptr(), C.byref(self.k), chunk,
                                                        self._stream(), C.byref(ex)), "md_rollout_graph_create")
            self._graph, self._graph_chunk, self._graph_w = ex, chunk, self.w
        if self._graph is not None and self._graph_chunk == chunk and self._graph_w is self.w:
            _lib.check(self.lib.md_rollout_graph_launch(self._graph, self._stream()), "md_rollout_graph_launch")
        else:
            _lib.check(self.lib.md_rollout_steps(self._g(), C.byref(s), self.w.ptr(), C.byref(self.k), chunk,
                                                 self._stream()), "md_rollout_steps")
        self.kernel_launches += chunk * ((5 if self.fused else 4 + _lib.BP_ITERS))
        if s is self.s_tail:
            _lib.check(self.lib.md_rollout_tail(self._g(), C.byref(s), self.w.ptr(), self._stream()), "md_rollout_tail")
            self.kernel_launches += 1

    def _finish_run(self):
        if self.tail:
            _lib.check(self.lib.md_tail_join(self._stream()), "md_tail_join")
        self._runs += 1

    @_on_device
    def run(self, chunk=8, step=1):
        """Greedy rollout to termination (GetSol when step == 1, GetSolution's top-`step` batches otherwise)."""
        B = self.batch.B
        it = 0
        topk = None
        if step > 1:
            topk = torch.empty(B, step, dtype=torch.int32, device=self.batch.device)
        self._begin_run(step)
        while True:
            if step == 1:
                self._enqueue_chunk(chunk, it == 0)
            else:
                for _ in range(chunk):
                    self.forward()
                    _lib.check(self.lib.md_select_topk(self._g(), C.byref(self.s), C.byref(self.k),
                                                       C.c_void_p(topk.data_ptr()), step, self._stream()), "md_select_topk")
                    self.kernel_launches += 1
                    for j in range(step):
                        self.step(topk, j)
            it += chunk
            if self.num_left() >= B or it > self.max_steps + chunk:
                break
        self._finish_run()
        return self.results()

    @_on_device
    def run_heuristic(self, kind="hda", combine="max", chunk=8):
        """critical_number of the reference's baseline scripts (U/baseline/HDA/hda_2max.py:71-93, CI/ci_max.py:92-110)
        for every graph of the batch: select by alive degree ("hda") or collective influence ("ci"), combined over the
        layers by "max" or "add", remove, recompute the LMCC, until LMCC == 1.  Needs no weights."""
        k, c = {"hda": 0, "ci": 1}[kind], {"max": 0, "add": 1}[combine]
        it = 0
        while True:
            for _ in range(chunk):
                _lib.check(self.lib.md_heuristic_select(self._g(), C.byref(self.s), k, c, self._stream()), "md_heuristic_select")
                self.step()
                self.kernel_launches += 1
            it += chunk
            if self.num_done() >= self.batch.B or it > self.max_steps + chunk:
                break
        return self.results()

    def __del__(self):
        try:
            if getattr(self, "_graph", None) is not None:
                self.lib.md_rollout_graph_destroy(self._graph)
        except Exception:
            pass

    def results(self):
        """Per-graph AUDC, number of removals, removal sequence, LMCC and MaxCC traces as numpy arrays: five asynchronous
        copies into pinned staging buffers (allocated once per rollout object) and ONE stream synchronize -- `.cpu()` goes
        through pageable memory and blocks the host once per tensor, which with several batches in flight keeps the other
        batches' next chunks from being enqueued."""
        src = dict(score=self.score, n_steps=self.n_steps, actions=self.trace_action, ranks=self.trace_rank, maxcc=self.trace_maxcc)
        pins = getattr(self, "_pinned_results", None)
        if pins is None:
            pins = self._pinned_results = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in src.items()}
        with torch.cuda.device(self.batch.device):
            for k, v in src.items():
                pins[k].copy_(v, non_blocking=True)
            torch.cuda.current_stream(self.batch.device).synchronize()
        return {k: p.numpy().copy() for k, p in pins.items()}

    # -- inspection (tests / Predict mirror) ----------------------------------------------------------
    def q_values(self, gi):
        """(active local ids ascending, q f32) of graph gi after forward()."""
        n = int(self.n_act[gi])
        S = self.batch.S
        ids = self.act_list[gi * S:gi * S + n].long()
        return (ids - gi * S).cpu().numpy(), self.q[ids].cpu().numpy()


class RolloutPipeline:
    """Software pipeline over a QUEUE of rollout jobs: up to `depth` jobs are in flight, each on its own CUDA stream.

    A single lock-step batch ends with a long under-occupied phase (the last few big graphs of the batch keep the
    iteration going while most SMs idle); with the next batch already in flight on another stream its full-width
    kernels fill those SMs.  Jobs never interact: results are identical to running them one after the other.

    make_job(j) is called with the job's stream current and returns the BatchRollout to run (a persistent object that
    the pipeline resets, or a freshly built one: GraphBatch upload + state allocation then happen on that stream);
    finish_job(j, rollout) is called -- in job order, with the job's stream current -- when job j is terminal
    (gather / download its results there)."""

    def __init__(self, device, depth=2, chunk=8):
        self.device = torch.device(device)
        self.depth, self.chunk = depth, chunk
        self.streams = [torch.cuda.Stream(self.device) for _ in range(depth)]
        self.events = [torch.cuda.Event() for _ in range(depth)]
        self._pins = [torch.zeros(1, dtype=torch.int32).pin_memory() for _ in range(depth)]

    def run(self, n_jobs, make_job, finish_job=None, reset=True):
        cur = torch.cuda.current_stream(self.device)
        for st in self.streams:
            st.wait_stream(cur)
        slots = [None] * self.depth  # slot -> [job index, rollout, iterations enqueued, terminal?]
        nxt, done_upto = 0, 0

        def enqueue(i):
            job = slots[i]
            with torch.cuda.stream(self.streams[i]):
                job[1]._enqueue_chunk(self.chunk, job[2] == 0)
                self._pins[i].copy_(job[1].flags[8:9], non_blocking=True)
                self.events[i].record(self.streams[i])
            job[2] += self.chunk

        while nxt < n_jobs or any(s is not None for s in slots):
            for i in range(self.depth):
                if slots[i] is None and nxt < n_jobs:
                    with torch.cuda.stream(self.streams[i]), torch.cuda.device(self.device):
                        ro = make_job(nxt)
                        if reset:
                            ro.reset()
                        ro._begin_run()
                    slots[i] = [nxt, ro, 0, False]
                    nxt += 1
                    enqueue(i)
            # advance every slot whose chunk has completed; block on the OLDEST job if none has
            order = sorted((i for i in range(self.depth) if slots[i] is not None), key=lambda i: slots[i][0])
            ready = [i for i in order if not slots[i][3] and self.events[i].query()]
            if not ready:
                waiting = [i for i in order if not slots[i][3]]
                if waiting:
                    self.events[waiting[0]].synchronize()
                    ready = [waiting[0]]
            for i in ready:
                job = slots[i]
                if int(self._pins[i][0]) >= job[1].batch.B or job[2] > job[1].max_steps + self.chunk:
                    job[3] = True
                else:
                    enqueue(i)
            # finalize terminal jobs in job order (keeps every rank's collectives in the same order)
            for i in order:
                job = slots[i]
                if job[3] and job[0] == done_upto:
                    with torch.cuda.stream(self.streams[i]), torch.cuda.device(self.device):
                        job[1]._finish_run()
                        if finish_job is not None:
                            finish_job(job[0], job[1])
                    slots[i] = None
                    done_upto += 1
        for st in self.streams:
            cur.wait_stream(st)


class MultiStreamRollout:
    """The same lock-step rollout, with the batch cut into `parts` independent sub-batches that advance on their own CUDA
    streams.  Graphs never interact, so this changes nothing in the results; what it buys is overlap: the env step of
    a sub-batch waits for its slowest graph and the tail of every kernel leaves SMs idle -- the other
    sub-batches' kernels fill them.  Sub-batch p holds graphs [p * ceil(B / parts), ...) in order.

    order="size": the graphs are first sorted by their number of edges (descending), so that graphs of similar size -- and
    hence rollouts of similar length -- share a sub-batch.  A lock-step batch runs as long as its longest rollout (C3: 390
    removals for a mean of 68): mixed with short ones the long rollouts keep a wide, mostly finished batch iterating,
    sorted they form a narrow sub-batch of their own whose iterations are cheap and whose tail the other sub-batches fill.
    results() always returns the graphs in their ORIGINAL order.  parts=None: one sub-batch per 128 graphs, 2..8.
    MEASURED on C3 (1024 unit-cost graphs, one GPU): 1 / 2 / 4 / 8 size-sorted sub-batches = 98 / 100 / 121 / 129 ms per pass --
    the pass is bound by its longest rollout's chain of ~2000 dependent launches, and every other sub-batch's kernels
    (persistent CTAs that own an SM for ~100 us) delay that chain more than they fill its idle SMs.  Sub-batches pay when
    there are many batches of similar length (RolloutPipeline: +70 % on C2), not inside one batch: bench.py and GetSolBatch
    keep a single lock-step batch per rank."""

    def __init__(self, graphs, weights, device=None, cost=None, parts=2, chunk=8, order=None):
        graphs = list(graphs)
        if parts is None:
            parts = max(2, min(8, len(graphs) // 128))
        parts = max(1, min(parts, len(graphs)))
        self.perm = np.arange(len(graphs))
        if order == "size":
            self.perm = np.argsort([-sum(int(e) for e in g.num_edges) for g in graphs], kind="stable")
        elif order is not None:
            raise ValueError("order: None or 'size'")
        graphs = [graphs[i] for i in self.perm]
        self.inv = np.argsort(self.perm, kind="stable")
        per = (len(graphs) + parts - 1) // parts
        self.slices = [graphs[i:i + per] for i in range(0, len(graphs), per)]
        self.batches = [GraphBatch(s, device, cost) for s in self.slices]
        self.rollouts = [BatchRollout(b, weights) for b in self.batches]
        self.device = self.batches[0].device
        self.streams = [torch.cuda.Stream(self.device) for _ in self.rollouts]
        self.events = [torch.cuda.Event() for _ in self.rollouts]
        self.chunk = chunk
        self._pins = [torch.zeros(1, dtype=torch.int32).pin_memory() for _ in self.rollouts]

    @property
    def kernel_launches(self):
        return sum(r.kernel_launches for r in self.rollouts)

    def run(self):
        cur = torch.cuda.current_stream(self.device)
        for ro, st in zip(self.rollouts, self.streams):
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                ro.reset()
                ro._begin_run()
        # every sub-batch advances on its own: a chunk is enqueued as soon as the previous one of the SAME sub-batch has
        # completed and left graphs alive (no round structure: a narrow sub-batch of long rollouts must not wait for the
        # chunks of the wide ones)
        its = [0] * len(self.rollouts)

        def enqueue(i):
            ro = self.rollouts[i]
            with torch.cuda.stream(self.streams[i]):
                ro._enqueue_chunk(self.chunk, its[i] == 0)
                self._pins[i].copy_(ro.flags[8:9], non_blocking=True)
                self.events[i].record(self.streams[i])
            its[i] += self.chunk

        live = list(range(len(self.rollouts)))
        for i in live:
            enqueue(i)
        while live:
            ready = [i for i in live if self.events[i].query()]
            if not ready:
                self.events[live[0]].synchronize()
                ready = [live[0]]
            for i in ready:
                ro = self.rollouts[i]
                if int(self._pins[i][0]) >= ro.batch.B or its[i] > ro.max_steps + self.chunk:
                    live.remove(i)
                else:
                    enqueue(i)
        for ro, st in zip(self.rollouts, self.streams):
            with torch.cuda.stream(st):
                ro._finish_run()
            cur.wait_stream(st)
        return self.results()

    def results(self):
        rs = [r.results() for r in self.rollouts]
        T = max(r["actions"].shape[1] for r in rs)
        pad = lambda a, fill: np.concatenate([a, np.full((a.shape[0], T - a.shape[1]), fill, a.dtype)], 1)
        inv = self.inv
        return dict(score=np.concatenate([r["score"] for r in rs])[inv], n_steps=np.concatenate([r["n_steps"] for r in rs])[inv],
                    actions=np.concatenate([pad(r["actions"], -1) for r in rs])[inv],
                    ranks=np.concatenate([pad(r["ranks"], 0) for r in rs])[inv],
                    maxcc=np.concatenate([pad(r["maxcc"], 0) for r in rs])[inv])
