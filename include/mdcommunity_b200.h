/*
 * mdcommunity_b200.h -- C ABI of the B200-native MultiDismantler rollout path (libmdb200.so, sm_100a only).
 *
 * The reference (KelvinRyman/MDCommunity) is pure Python and has no FFI of its own; the entry points below
 * are what a binding for its hot path would call.  Each one names the reference code it replaces
 * (paths relative to /root/reference/code; U/ = MultiDismantler_unit_cost/, D/ = MultiDismantler_degree_cost/).
 *
 * Conventions (all entry points):
 *   - extern "C", plain pointers and sizes only; every pointer inside the structs is a DEVICE pointer unless
 *     the field comment says "host".
 *   - return 0 on success, a negative MD_E* code for bad arguments, a positive cudaError_t for CUDA failures.
 *   - nothing is allocated, nothing synchronises: the caller owns all buffers (sized with the helpers at the end)
 *     and orders work with the cudaStream_t it passes.
 *   - one host thread per GPU process.  There is no CPU fallback.
 *
 * Node addressing: a batch holds B graphs; graph g owns the global node slots [g*S, (g+1)*S) where the stride
 * S is a multiple of MD_TILE_M and S > n_nodes[g] for every graph (at least one padding slot: the alive-CSR row pointers
 * of graph g occupy aptr[g*S .. g*S + n_act[g]], which must not reach the first word of graph g + 1).
 * Slots >= n_nodes[g] are padding (label -1 for ever).
 */
#ifndef MDCOMMUNITY_B200_H
#define MDCOMMUNITY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MD_MAX_LAYERS 4
#define MD_EMBED 64     /* EMBEDDING_SIZE, U/MultiDismantler_torch.py:35 */
#define MD_HIDDEN 32    /* REG_HIDDEN, :53 */
#define MD_AUX 4        /* aux_dim, :58 */
#define MD_GATE 128     /* w_layer1 width, U/MultiDismantler_net_graphsage.py:83 */
#define MD_BP_ITERS 3   /* max_bp_iter, U/MultiDismantler_torch.py:62 */
#define MD_TILE_M 128   /* node rows per CTA tile in the embedding / Q kernels */

#define MD_EINVAL (-1)
#define MD_ETOOBIG (-2)
#define MD_ELAYERS (-3)

typedef void* md_stream_t; /* cudaStream_t */

/* Static topology of a batch (replaces graph.Graph / Graph_test adjacency lists, U/graph.py:8-30,70-84,
 * and the per-step PrepareBatchGraph COO rebuild, U/PrepareBatchGraph.py:76-177: the CSR is built once). */
typedef struct md_graphs {
  int32_t n_graphs;                      /* B */
  int32_t n_layers;                      /* L: 2 for the reference network; 3 allowed for the LMCC / aggregation kernels */
  int32_t stride;                        /* S, multiple of MD_TILE_M */
  int32_t n_total;                       /* B * S */
  int32_t max_graph_nnz;                 /* max over graphs of the CSR entries of one graph summed over layers (host) */
  int32_t reserved0;
  const int32_t* rowptr[MD_MAX_LAYERS];  /* [n_total + 1] per layer */
  const int32_t* col[MD_MAX_LAYERS];     /* [nnz_l] neighbour = global node id; row order = edge_list order */
  const int32_t* src[MD_MAX_LAYERS];     /* [nnz_l] row (global node id) of every CSR entry, i.e. the COO row array */
  const int32_t* n_nodes;                /* [B] */
  const int32_t* n_edges;                /* [L][B] undirected edge counts (graph.num_edges) */
  const int32_t* max_rank;               /* [B] LMCC of the intact graph (graph.max_rank) */
  const float* node_w[MD_MAX_LAYERS];    /* degree cost: [n_total] w_l(v) = deg/maxdeg as f32 (D/graph.py:39-51); NULL for unit cost */
  const double* node_cost;               /* degree cost: [n_total] f64 w0[v]/W0 + w1[v]/W1 (D/mvc_env.py:127-133); NULL for unit cost */
} md_graphs_t;

/* Mutable environment state of the batch (replaces MvcEnv fields covered_set / remove_edge / numCoveredEdges /
 * score / MaxCCList, U/mvc_env.py:8-52, and PrepareBatchGraph.get_status_info's counters, U/PrepareBatchGraph.py:35-74). */
typedef struct md_state {
  int32_t* label;        /* [n_total] mutual-component id = global id of its smallest node; -1 = removed or padding */
  int32_t* deg[MD_MAX_LAYERS]; /* [n_total] alive degree per layer */
  int32_t* act_list;     /* [n_total] graph g's active nodes, ascending, at [g*S, g*S + n_act[g]) */
  int32_t* n_act;        /* [B] */
  int32_t* aptr[MD_MAX_LAYERS]; /* [n_total + 1] CSR of the ALIVE edges in compact order: row g*S + p is active node p of graph g */
  int32_t* acol[MD_MAX_LAYERS]; /* [nnz_l] alive neighbours (global ids); graph g uses its static CSR region */
  int32_t* asrc[MD_MAX_LAYERS]; /* [nnz_l] row (global node id) of every alive entry: the next env step starts from the alive CSR */
  int32_t* n_removed;    /* [B] |covered_set| */
  int32_t* cov_edges;    /* [L][B] numCoveredEdges */
  int32_t* alive_edges;  /* [L][B] */
  int32_t* max_deg;      /* [L][B] */
  int64_t* twohop;       /* [L][B] sum_v C(deg_alive(v), 2) */
  int32_t* rank;         /* [B] current LMCC size */
  int32_t* done;         /* [B] isTerminal */
  int32_t* n_steps;      /* [B] removals so far */
  double* score;         /* [B] AUDC accumulator (f64, reference op order) */
  float* aux;            /* [B][L][4] aux features for the next forward */
  unsigned long long* best; /* [B] packed (orderable q << 32 | ~local id) arg-max key, written by md_qhead */
  int32_t* trace_action; /* [B][max_steps] local node ids removed */
  int32_t* trace_rank;   /* [B][max_steps] LMCC after each removal */
  double* trace_maxcc;   /* [B][max_steps] MaxCCList entries */
  int32_t* n_done;       /* [1] number of terminal graphs */
  int32_t max_steps;
  int32_t cost_mode;     /* 0 = unit cost reward, 1 = degree cost reward */
  /* scratch for the large-graph (multi-CTA) LMCC path, may be NULL when stride <= md_env_small_max_stride() */
  int32_t* parent;       /* [n_total] */
  int32_t* comp_size;    /* [n_total]; [3 * n_total] when lg_scratch is set */
  int32_t* flags;        /* [32] scratch: [0..3] grid barrier / change flags of the multi-CTA env kernel, [4..5] tile queue of
                            md_qhead, [6..7] phase timers of the multi-CTA env kernel, [8] number of graphs that have LEFT the
                            lock-step iteration (terminal, or handed over to md_rollout_tail), [9..21] counters of the multi-CTA env kernel's
                            incremental update ([16] = next fresh label, persists between launches) */
  int32_t* cur_action;   /* [B] action applied in the running step (-1 = graph untouched) */
  int32_t* tail_n;       /* [B] or NULL.  Non-NULL enables the hand-over to md_rollout_tail: when a step leaves a graph with
                            at most md_tail_max_nodes() active nodes (and md_tail_max_entries() alive entries per layer),
                            md_env_step stores that node count here and marks the graph finished FOR THE LOCK-STEP KERNELS
                            (n_act = 0, done = 1, n_done untouched); md_rollout_tail completes it.  Zero-initialised. */
  int32_t* lg_scratch;   /* [9 * n_total] or NULL, ZERO-INITIALISED: scratch of the incremental mutual-component update of the
                            multi-CTA env kernel (NULL = from-scratch prune rounds on every step); the kernel leaves its first 2 * n_total words zero */
} md_state_t;

/* The 14-tensor checkpoint (U/MultiDismantler_net_graphsage.py:42-89, U/MRGNN/mutil_layer_weight.py:252-264),
 * row-major f32 exactly as stored, plus two host-precomputed products used by the fused iteration kernel. */
typedef struct md_weights {
  const float* w_n2l;     /* [2][64] */
  const float* conv1;     /* [64][64]  p_node_conv */
  const float* conv2;     /* [64][64]  p_node_conv2 */
  const float* conv3;     /* [128][64] p_node_conv3 */
  const float* w13;       /* [64][64]  p_node_conv  @ p_node_conv3[:64]  (formed in f64, rounded once) */
  const float* w23;       /* [64][64]  p_node_conv2 @ p_node_conv3[64:] */
  const float* trans;     /* [64][64]  layerNodeAttention_weight.trans */
  const float* bias;      /* [64] */
  const float* logis_w;   /* [64] */
  const float* logis_b;   /* [1] */
  const float* cross;     /* [64]  cross_product */
  const float* h1;        /* [64][32] */
  const float* h2;        /* [36] (= last_w) */
  const float* wl1;       /* [64][128] w_layer1 */
  const float* wl2;       /* [128] w_layer2 */
} md_weights_t;

/* Per-step activations. */
typedef struct md_work {
  float* h[2][MD_MAX_LAYERS];   /* ping-pong node embeddings, [n_total][64] each */
  float* psum;                  /* [MD_BP_ITERS][L][n_total / MD_TILE_M][64] per-tile partial sums of H (subgsum) */
  float* gconst;                /* [B][MD_GCONST] per-graph constants for the Q head */
  float* q;                     /* [n_total] Q value per node slot (valid for active nodes) */
  void* fused_ws;               /* md_fused_ws_bytes(g) bytes, 1024-byte aligned: scratch of md_embed_fused (weight images,
                                   unit counter, per-SM L2-resident column-half images); NULL = resident path not used */
} md_work_t;
#define MD_GCONST 16 /* s[L], gate_softmax[L], aux_dot[L] padded */

/* ---- environment ------------------------------------------------------------------------------------- */

/* MvcEnv.s0 (U/mvc_env.py:31-52) for every graph of the batch: clears the state and runs the initial
 * mutual-component prune (Mcc.MCC, U/Mcc.py:30-38). */
int md_env_reset(const md_graphs_t* g, const md_state_t* s, md_stream_t stream);

/* Arbitrary environment state for Predict on replay states (U/MultiDismantler_torch.py:263-302 called from Fit :315-344):
 * removed_mask[n_total] (1 = node in covered_set); the caller writes numCoveredEdges into s->cov_edges first
 * (the reference recounts it from remove_edge in get_status_info, U/PrepareBatchGraph.py:45-52).  Labels, degrees,
 * the alive CSR and the aux features are recomputed; score / traces are cleared. */
int md_env_set_state(const md_graphs_t* g, const md_state_t* s, const uint8_t* removed_mask, md_stream_t stream);

/* MvcEnv.stepWithoutReward (U/mvc_env.py:74-87) + getReward / getMaxConnectedNodesNum (:133-162) + Mcc.MCC for
 * every non-terminal graph, followed by PrepareBatchGraph.get_status_info (U/PrepareBatchGraph.py:35-100).
 * action source: actions != NULL -> actions[g * action_stride + (action_index >= 0 ? action_index : n_steps[g])]
 * (local node id, < 0 = skip this graph); actions == NULL -> the arg-max key md_qhead left in s->best. */
int md_env_step(const md_graphs_t* g, const md_state_t* s, const int32_t* actions, int32_t action_stride,
                int32_t action_index, md_stream_t stream);

/* largest graph stride the single-CTA (shared-memory) env kernel handles; larger strides use the multi-CTA path */
int32_t md_env_small_max_stride(void);
int32_t md_env_uses_small_path(const md_graphs_t* g);
/* test hook: route every batch through the multi-CTA (global-memory, cooperative) env kernel */
void md_debug_force_large_env(int32_t on);
/* profiling hook: bit 0 skips the neighbour gather, bit 1 the dense transform of md_embed_iter (results become wrong);
 * bit 2 makes md_rollout_steps use the unfused embedding kernels even where md_embed_fused applies;
 * bit 3 makes md_qhead use its FP32-FMA kernel instead of the tensor-core kernel;
 * bit 4 makes md_embed_iter use its single-kernel tensor-core variant (measured no faster than the FP32-FMA kernel);
 * bit 6 switches off the L2 prefetch of the fused md_embed_iter (embeddings beyond L2);
 * bit 7 disables / bit 8 forces the two-kernel md_embed_iter (md_agg neighbour sums + tensor-core transform; default
 *   only when the embeddings exceed L2), bit 9 selects md_agg's register-load kernel instead of the cp.async ring,
 *   bit 12 disables iteration 0 from the alive degrees (unit cost) so that md_embed_init writes H0 again;
 * bit 18 (two-kernel md_embed_iter, unit cost): materialise H1 -- by default iteration 0 launches nothing and iteration 1
 *   takes both its neighbour sums and its own rows from the table H1(alive degree) (see md_embed_iter);
 * bit 13 keeps relu's argument e @ h1 of the two-layer tensor-core md_qhead on the FP32 pipes (A/B reference);
 * bit 17 (md_env_step, multi-CTA path): from-scratch prune rounds on every step instead of the grid-wide incremental
 *   mutual-component update that is the default when s->lg_scratch is set (A/B reference; same results);
 * bit 16 (md_env_step, single-CTA path): incremental mutual-component update (multi-source search from the end points of
 *   the edges that died) instead of the from-scratch prune rounds; same results, measured slower at N = 1000;
 * bits 14 / 15 (two-kernel md_embed_iter only): enqueue only the neighbour aggregation / only the transform, so that a
 *   caller can time the two kernels of an iteration separately (bit 14 call followed by bit 15 call == the normal sequence);
 * bits 10 / 11 (timing experiments, results become wrong / unchanged): transform kernel without its global stores /
 *   without its L2 prefetch of the next tile;
 * bit 19 (md_embed_fused): every (graph, layer) unit alone and on the single-unit code path, i.e. the kernel as it was
 *   before small units were packed (A/B reference: embeddings within rounding of the column sums' order) */
void md_debug_set_flags(int32_t flags);
/* md_embed_fused packs (graph, layer) units of few active rows into bins that one persistent CTA processes together.
 * rows = 0: weight budget of a bin chosen per launch (the small units' total spread over the SMs, whole tiles, 256..768);
 * rows > 0: that budget (clamped to 256..768); rows < 0: every unit alone.  Results do not depend on the setting. */
void md_debug_set_fused_pack(int32_t rows);
/* profiling hook: clock64 phase times of CTA 0 of the last md_env_step (labels in, CSR staged, prune rounds,
 * component sizes, alive CSR + degrees) and, at index 16, the number of prune rounds */
int md_debug_read_env(long long* host, int n);
/* profiling hook: per graph (first 1024) of the last md_env_step: {cycles, prune rounds, staged CSR entries, SM id} */
int md_debug_read_env_ctas(unsigned* host4096);
/* profiling hook: clock64 phase sums of CTA 0 of the last md_embed_fused launch.  Gather warp 0: [0] unit setup,
 * [1] first-embedding fill, [2] column-half reload, [3] gather, [4] wait for the staging tile, [5] end-of-pass wait,
 * [6] epilogue, [7] idle between units; MMA thread: [8] weight-image wait, [9] wait for a full tile, [10] MMA issue,
 * [11] end-of-pass completion wait; [15] units processed */
int md_debug_read_fused(long long* host, int n);
/* profiling hook: per CTA (up to 160) of the last md_embed_fused launch: {cycles in the unit loop, units processed} */
int md_debug_read_fused_ctas(unsigned* host320);
/* profiling hook: gather sub-phase cycles of CTA 0 / warp 0: row setup, neighbour loop, own rows, staging stores,
 * long rows, proxy fence + arrive */
int md_debug_read_fused_gather(unsigned* host8);

/* ---- Q network ---------------------------------------------------------------------------------------- */

/* Node input features + first embedding (U/MultiDismantler_net_graphsage.py:248-287; D/MultiDismantler_net.py:222-240)
 * into h[0].  A no-op for unit-cost batches whose embeddings exceed L2: there md_embed_iter(iter = 0) works from the
 * alive degrees and never reads h[0] (call the two in sequence as md_rollout_steps does). */
int md_embed_init(const md_graphs_t* g, const md_state_t* s, const md_weights_t* w, const md_work_t* k,
                  md_stream_t stream);

/* One message-passing iteration for all layers: neighbour sum over alive edges (torch_sparse.spmm n2n, :290-291),
 * dense transforms (:294-309), relu + L2 normalise (:309-310), and the per-graph sum of the OLD embeddings
 * (spmm subgsum, :297-298) as per-tile partials.  iter = 0..2; reads h[iter&1], writes h[(iter+1)&1].
 * Embeddings beyond L2 (N = 1M): two launches, md_agg's neighbour sums into h[(iter+1)&1] and a tensor-core transform
 * that overwrites them in place.  md_graph_embed consumes the per-tile partials (call it once per forward).
 * Unit-cost batches on that path: the features are [d/dmax, d/dmax], so relu(x @ w_n2l) normalised is the SAME vector c
 * for every node with an alive neighbour, the first neighbour sum is deg(v) c, and the embedding after iteration 0 is
 * H1(d) = normalise(relu(d (c @ w13) + c @ w23)), a function of the alive degree d only.  iter = 0 therefore enqueues
 * nothing and leaves h[1] untouched; iter = 1 gathers 4-byte degrees instead of 256-byte rows and also writes iteration
 * 0's per-tile partials; from iter = 2 on everything is as described above (md_debug_set_flags bit 18 materialises H1). */
int md_embed_iter(const md_graphs_t* g, const md_state_t* s, const md_weights_t* w, const md_work_t* k,
                  int32_t iter, md_stream_t stream);

/* md_embed_init + the MD_BP_ITERS md_embed_iter calls in ONE launch for batches that satisfy md_fused_applies():
 * a persistent CTA per (graph, layer) keeps the embeddings resident in shared memory (one 32-column half at a time),
 * gathers neighbour rows from there and runs the dense transforms on the tensor cores (tcgen05, 3-term tf32/bf16
 * split, accumulators in tensor memory).  Same outputs as the unfused sequence: h[MD_BP_ITERS & 1] and psum.
 * Returns MD_ETOOBIG for larger strides and MD_EINVAL when k->fused_ws is NULL. */
int md_embed_fused(const md_graphs_t* g, const md_state_t* s, const md_weights_t* w, const md_work_t* k,
                   md_stream_t stream);
int32_t md_fused_max_stride(void);
/* 1 if md_embed_fused handles this batch: stride <= md_fused_max_stride() and < 65536 CSR entries per graph */
int32_t md_fused_applies(const md_graphs_t* g);
int64_t md_fused_ws_bytes(const md_graphs_t* g);

/* Virtual (graph) node chain y0..y3 (:272-287,301-321), its inter-layer fusion rows
 * (BitwiseMultipyLogis, U/MRGNN/mutil_layer_weight.py:266-313) and the per-graph Q-head constants. */
int md_graph_embed(const md_graphs_t* g, const md_state_t* s, const md_weights_t* w, const md_work_t* k,
                   md_stream_t stream);

/* Inter-layer fusion of node rows + Q head + masked arg-max (net_graphsage.py:323-393, Predict's -inf mask
 * U/MultiDismantler_torch.py:286-300 and argMax :799-807).  Reads h[MD_BP_ITERS & 1]. */
int md_qhead(const md_graphs_t* g, const md_state_t* s, const md_weights_t* w, const md_work_t* k,
             md_stream_t stream);

/* n_steps lock-step greedy iterations (GetSol loop body, U/MultiDismantler_torch.py:759-784): forward + env step. */
int md_rollout_steps(const md_graphs_t* g, const md_state_t* s, const md_weights_t* w, const md_work_t* k,
                     int32_t n_steps, md_stream_t stream);

/* ---- tail of a rollout ------------------------------------------------------------------------------------------
 * A lock-step iteration costs ~100 us of fixed latency however small the graphs have become, and most removals of a
 * rollout happen after the graph has collapsed to a few dozen active nodes.  With s->tail_n set, such graphs leave the
 * lock step (see md_state_t.tail_n) and ONE CTA per graph finishes the rollout on chip: the GetSol loop body
 * (U/MultiDismantler_torch.py:759-784: Predict + arg-max + stepWithoutReward) with all state in shared memory.
 * Two-layer batches on the single-CTA env path only (md_tail_applies). */
int32_t md_tail_applies(const md_graphs_t* g);
int32_t md_tail_max_nodes(void);
int32_t md_tail_max_entries(void);
/* hand over every non-terminal graph that is already small enough (call once after md_env_reset, before the first iteration) */
int md_tail_handover(const md_graphs_t* g, const md_state_t* s, md_stream_t stream);
/* Finish all handed-over graphs.  The launch is ordered AFTER everything enqueued on `stream` so far but runs on an
 * internal low-priority stream, i.e. it does not block later work on `stream`: the lock-step iterations of the larger
 * graphs continue while the small ones finish on idle SMs.  md_tail_join makes `stream` wait for all tail launches. */
int md_rollout_tail(const md_graphs_t* g, const md_state_t* s, const md_weights_t* w, md_stream_t stream);
int md_tail_join(md_stream_t stream);
/* profiling hook (profiling builds only, zeros otherwise): per graph (first 1024) of the tail launches so far
 * {cycles, steps, nodes at hand-over, SM id, phase cycles: setup, embedding iterations, graph rows, Q head + arg-max,
 * LMCC rounds, rest of the env step} */
int md_debug_read_tail(unsigned long long* host10240);

/* The same n_steps iterations captured once into a CUDA graph and replayed (launch-bound tail of a rollout).  Call
 * md_rollout_steps at least once before md_rollout_graph_create; the graph bakes in the struct values (device pointers),
 * so it is valid while the caller keeps the same buffers.  Small-graph (single-CTA env) batches only: MD_ETOOBIG else. */
int md_rollout_graph_create(const md_graphs_t* g, const md_state_t* s, const md_weights_t* w, const md_work_t* k,
                            int32_t n_steps, md_stream_t stream, void** exec_out);
int md_rollout_graph_launch(void* exec, md_stream_t stream);
int md_rollout_graph_destroy(void* exec);

/* Top-k selection for GetSolution's stepRatio mode (np.argsort(-pred)[:step], U/MultiDismantler_torch.py:725):
 * out_local_ids[g][0..topk) = graph g's active nodes ranked by (q descending, id ascending), -1 padded.
 * Feed it to md_env_step(actions = out_local_ids, action_stride = topk, action_index = j) for j = 0..topk-1. */
int md_select_topk(const md_graphs_t* g, const md_state_t* s, const md_work_t* k, int32_t* out_local_ids,
                   int32_t topk, md_stream_t stream);

/* ---- heuristic baselines ------------------------------------------------------------------------------ */

/* Selection step of the heuristic baselines (U/baseline/HDA/hda_2max.py:50-69, hda_add.py; U/baseline/CI/ci_max.py:52-90,
 * ci_add.py) on the current environment state: kind 0 = HDA (alive degree), 1 = CI (radius-1 collective influence);
 * combine 0 = max over layers, 1 = sum.  Leaves the arg-max key (lowest id among the maxima) in s->best, exactly as
 * md_qhead does, so md_env_step(actions = NULL) applies it: one md_heuristic_select + one md_env_step = one iteration of
 * critical_number (hda_2max.py:71-93). */
int md_heuristic_select(const md_graphs_t* g, const md_state_t* s, int32_t kind, int32_t combine, md_stream_t stream);

/* ---- operator boundary -------------------------------------------------------------------------------- */

/* torch_sparse.spmm replacement (call sites net_graphsage.py:290,297,350,377): Y[r] = sum_e val[e] * X[col[e]]
 * over CSR row r.  val == NULL means all ones.  D must be a multiple of 4, X and Y 16-byte aligned. */
int md_spmm_csr_f32(const int32_t* rowptr, const int32_t* col, const float* val, const float* X, float* Y,
                    int32_t n_rows, int32_t D, md_stream_t stream);

/* ---- dense primitives of the DQN training step, FP32, forward and backward -------------------
 * Replace the torch ops of MultiDismantler_net.train_forward (U/MultiDismantler_net_graphsage.py:92-241), of
 * BitwiseMultipyLogis (U/MRGNN/mutil_layer_weight.py:266-313), of calc_loss (U/MultiDismantler_torch.py:410-431), the
 * backward torch autograd derives from them, and torch.optim.Adam (U/MultiDismantler_torch.py:136).  All matrices are
 * dense row-major f32 device pointers; every reduction runs in a fixed order (no floating-point atomics). */

/* C[M,N] = [C +] op(A) op(B); transA: A stored [K,M]; transB: B stored [N,K]; relu_a: A clamped at 0 on load.
 * ws / ws_floats: optional workspace for a split over K (used when K >= 1024 and the output has < 64 tiles). */
int md_gemm_f32(const float* A, const float* B, float* C, int32_t M, int32_t N, int32_t K, int32_t transA, int32_t transB,
                int32_t accumulate, int32_t relu_a, float* ws, int64_t ws_floats, md_stream_t stream);
/* H = x / max(|x|_2, 1e-12) per row, x = relu ? max(Z, 0) : Z  (F.normalize(torch.relu(.)), :113-116,:160-166); D <= 128 */
int md_norm_fwd_f32(const float* Z, float* H, int32_t n, int32_t D, int32_t relu, md_stream_t stream);
int md_norm_bwd_f32(const float* Z, const float* dH, float* dZ, int32_t n, int32_t D, int32_t relu, md_stream_t stream);
/* T = tanh(Z + b) (mutil_layer_weight.py:270-276); dZ = dT (1 - T^2) */
int md_tanh_bias_fwd_f32(const float* Z, const float* b, float* T, int32_t n, int32_t D, md_stream_t stream);
int md_tanh_bwd_f32(const float* T, const float* dT, float* dZ, int32_t n, int32_t D, md_stream_t stream);
/* out[c] = sum_r X[r, c]; ws / ws_floats: optional workspace (row chunks summed in parallel, folded in chunk order) */
int md_colsum_f32(const float* X, float* out, int32_t n, int32_t D, float* ws, int64_t ws_floats, md_stream_t stream);
/* two-layer fusion (mutil_layer_weight.py:278-301), width MD_EMBED: X_p = T_p + softmax_k(sigmoid(<T_k T_p, lw> + lb))[1-p] T_{1-p};
 * saved [n, 8] keeps the gate scalars for the backward; Glw [n, 64] / Glb [n] are per-row contributions to d lw / d lb */
int md_fuse2_fwd_f32(const float* T0, const float* T1, const float* lw, const float* lb, float* X0, float* X1, float* saved,
                     int32_t n, md_stream_t stream);
int md_fuse2_bwd_f32(const float* T0, const float* T1, const float* lw, const float* saved, const float* dX0, const float* dX1,
                     float* dT0, float* dT1, float* Glw, float* Glb, int32_t n, md_stream_t stream);
/* E[r,:] = H[r,:] * s[gid[r]]  ((hf yf^T) cross == hf * (yf . cross), net_graphsage.py:205-212) */
int md_rowscale_f32(const float* H, const float* s, const int32_t* gid, float* E, int32_t n, int32_t D, md_stream_t stream);
/* out[r] = A[r,:] . B[r,:] */
int md_rowdot_f32(const float* A, const float* B, float* out, int32_t n, int32_t D, md_stream_t stream);
/* out[g] = sum_{r in [off[g], off[g+1])} x[r] * (y ? y[r] : 1) */
int md_seg_reduce_f32(const float* x, const float* y, const int64_t* off, float* out, int32_t B, md_stream_t stream);
/* q[r] = sm0 (q0[r] + a0[g]) + sm1 (q1[r] + a1[g]), sm = softmax(gate0[g], gate1[g]), g = gid[r] (net_graphsage.py:214-241) */
int md_head_mix_fwd_f32(const float* g0, const float* g1, const float* a0, const float* a1, const float* q0, const float* q1,
                        const int32_t* gid, float* q, int32_t n, md_stream_t stream);
/* dsum_q[g] = sum_r dq (q0 - q1), dqs[g] = sum_r dq (md_seg_reduce_f32) */
int md_head_mix_bwd_f32(const float* g0, const float* g1, const float* a0, const float* a1, const float* dq, const float* dsum_q,
                        const float* dqs, const int32_t* gid, float* dq0, float* dq1, float* dg0, float* dg1, float* da0,
                        float* da1, int32_t n, int32_t B, md_stream_t stream);
/* out = alpha a + beta b (b may be NULL); dX *= (X > 0); out[0] = a . b */
int md_axpby_f32(const float* a, const float* b, float alpha, float beta, float* out, int64_t n, md_stream_t stream);
int md_relu_mask_f32(const float* X, float* dX, int64_t n, md_stream_t stream);
int md_dot_f32(const float* a, const float* b, float* out, int64_t n, md_stream_t stream);
/* one Adam step (t = 1, 2, ...) on flat parameter / gradient / moment buffers */
int md_adam_step_f32(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                     int32_t t, md_stream_t stream);

/* ---- sizing helpers (host) ------------------------------------------------------------------------------ */
int64_t md_psum_floats(int32_t n_layers, int32_t n_total);
/* sizeof(md_graphs_t), sizeof(md_state_t), sizeof(md_weights_t), sizeof(md_work_t) as compiled: lets a binding check
 * its own struct mirrors */
void md_abi_sizes(int32_t* out4);
const char* md_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MDCOMMUNITY_B200_H */
