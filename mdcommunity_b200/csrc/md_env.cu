// Environment step for the batch: node removal, LMCC recompute (iterative mutual-component pruning with a
// shared-memory union-find), AUDC accumulation and the per-graph status counters the next forward needs.
//
// Replaces (reference, /root/reference/code/MultiDismantler_unit_cost): mvc_env.py:31-52 (s0), :74-87
// (stepWithoutReward), :133-162 (getReward, getMaxConnectedNodesNum), Mcc.py:3-38 (MCC / deledge),
// PrepareBatchGraph.py:35-100 (get_status_info + aux features); degree-cost reward: D/mvc_env.py:127-133.
//
// State equivalence: the reference keeps (covered_set, remove_edge[l]).  Here a node's label is the id of its
// mutual component (smallest member); an edge is alive iff both end labels are equal and >= 0.  remove_edge is
// the set of edges between differently-labelled live nodes, so it never needs to be stored.
#include "md_common.cuh"

namespace md {
// phase clocks (profiling builds only: MDB200_PROFILE=1 python -m mdcommunity_b200.build --force; release builds carry
// no clock reads and no debug globals, so concurrent launches on several streams share no written state)
__device__ long long g_env_dbg[32];
__device__ unsigned g_env_cta[1024][4];  // per graph of the last step: cycles, prune rounds, staged entries, SM id
#ifdef MD_PROFILE
#define ENV_T0() long long t_dbg = clock64(); int k_dbg = 0
#define ENV_TICK() do { if (blockIdx.x == 0 && tid == 0) { long long n_ = clock64(); g_env_dbg[k_dbg] = n_ - t_dbg; t_dbg = n_; } ++k_dbg; } while (0)
#define ENV_CLK() clock64()
#define ENV_PROF(...) __VA_ARGS__
#else
#define ENV_T0() do { } while (0)
#define ENV_TICK() do { } while (0)
#define ENV_CLK() 0ll
#define ENV_PROF(...)
#endif

constexpr int ENV_T = 512;
constexpr int ENV_W = ENV_T / 32;
constexpr int ENV_SMALL_MAX = 16384;           // largest stride the single-CTA kernel accepts
constexpr size_t ENV_SMALL_SMEM = 200 * 1024;  // labels + union-find + staged CSR must fit in this much shared memory

__device__ __forceinline__ int uf_find(volatile int* par, int x) {
  int p = par[x];
  while (p != x) {  // path halving; parents only ever move towards smaller ids, racy writes stay valid
    int gp = par[p];
    if (gp != p) par[x] = gp;  // (p is a root: x already points at it, no store)
    x = gp;
    p = par[x];
  }
  return x;
}

__device__ __forceinline__ void uf_unite(int* par, int a, int b) {
  int ra = uf_find(par, a), rb = uf_find(par, b);
  while (ra != rb) {
    int hi = max(ra, rb), lo = min(ra, rb);
    int old = atomicCAS(&par[hi], hi, lo);  // hook the larger root under the smaller: roots are minima
    if (old == hi) break;
    ra = uf_find(par, old);
    rb = uf_find(par, lo);
  }
}

struct BlockRed {
  int i[ENV_W];
  long long ll[ENV_W];
};

__device__ __forceinline__ int block_sum_i(int v, BlockRed& r) {
  v = warp_sum_i(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) r.i[threadIdx.x >> 5] = v;
  __syncthreads();
  int t = 0;
#pragma unroll
  for (int k = 0; k < ENV_W; ++k) t += r.i[k];
  return t;
}
__device__ __forceinline__ int block_max_i(int v, BlockRed& r) {
  v = warp_max_i(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) r.i[threadIdx.x >> 5] = v;
  __syncthreads();
  int t = r.i[0];
#pragma unroll
  for (int k = 1; k < ENV_W; ++k) t = max(t, r.i[k]);
  return t;
}
__device__ __forceinline__ long long block_sum_ll(long long v, BlockRed& r) {
  v = warp_sum_ll(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) r.ll[threadIdx.x >> 5] = v;
  __syncthreads();
  long long t = 0;
#pragma unroll
  for (int k = 0; k < ENV_W; ++k) t += r.ll[k];
  return t;
}

// One CTA per graph.  init != 0: MvcEnv.s0; otherwise one stepWithoutReward.
// The graph's CSR (all layers) is copied into shared memory once, with graph-local node ids, together with the
// source row of every entry, so that (a) the ~10 edge sweeps of a step never leave the SM and (b) the union-find
// hooking runs edge-parallel (one thread per CSR entry) and hubs cannot serialise a warp.
// Besides the labels / degrees / counters the kernel emits, per layer, the CSR of the *alive* edges of the active
// nodes in compact (act_list) order: aptr[l][g*S + p] .. aptr[l][g*S + p + 1] indexes acol[l].  The three
// message-passing iterations of the next forward stream it without re-checking labels.
// IdxT: uint16_t when every graph of the batch has fewer than 65536 CSR entries (halves the staged CSR, so two CTAs fit
// one SM and a 256-graph batch runs in one wave instead of two), int32_t otherwise.
// INC: the incremental mutual-component update is compiled in (md_debug_set_flags bit 16; + 7 shared arrays per CTA).
template <typename IdxT, bool INC>
__global__ void __launch_bounds__(ENV_T) env_step_small(md_graphs_t G, md_state_t S, const int32_t* forced,
                                                       int forced_stride, int forced_index, int init, int max_nnz) {
  extern __shared__ int smem[];
  const int g = blockIdx.x;
  const int tid = threadIdx.x;
  const int lane = tid & 31, wid = tid >> 5;
  const int St = G.stride;
  const int base = g * St;
  const int N = G.n_nodes[g];
  const int L = G.n_layers;
  int* lab = smem;                  // [St] graph-local label, -1 = removed / padding   (lab and tmp swap roles after every
  int* par = smem + St;             // [St] union-find parent; later the alive degree     prune round: the flattened labels
  int* tmp = smem + 2 * St;         // [St] scratch; later the compact list of active nodes   are written into the other one)
  int* arow = smem + 3 * St;        // [St] node (graph-local id) of input row p
  int* c_rp = smem + 4 * St;        // [L][St + 1] row offsets into c_cl / c_src
  int* rowof = c_rp + L * (St + 1); // [St] node -> input row (-1: the node has no row)
  int* vis = rowof + St;            // [St] incremental LMCC: visited list (its tail segment is the current frontier)
  int* pend = vis + St;             // [St] pending-source bits, one per layer;  vis + pend = component-size table [2 St]
  int* srcl = pend + St;            // [St] pending sources of the running round
  int* gact = srcl + St;            // [St] per source group: has a frontier;     srcl + gact = smallest-member table [2 St]
  int* grmin = gact + St;           // [St] per source group: smallest node
  int* gnew = grmin + St;           // [St] per source group: fresh label (also scratch)
  // (the seven arrays above exist only when the incremental update is switched on: 28 KB at stride 1024 would cost the
  // second resident CTA per SM)
  constexpr bool inc_layout = INC;
  IdxT* c_cl = reinterpret_cast<IdxT*>(inc_layout ? gnew + St : rowof);  // [max_nnz] neighbour (graph-local id)
  IdxT* c_src = c_cl + max_nnz;     // [max_nnz] row of the entry (static, staged); later the alive prefix
  __shared__ BlockRed red;
  __shared__ int s_action, s_lab_a;
  __shared__ int s_scan[ENV_W];
  __shared__ int s_cnt[8];

  if (!init && S.done[g]) return;
  ENV_T0();
  const long long t_cta0 = ENV_CLK();
  (void)t_cta0;

  // ---- pick the action -----------------------------------------------------------------------------
  if (tid == 0) {
    int a = -1;
    if (!init) {
      if (forced != nullptr) {
        const int idx = forced_index >= 0 ? forced_index : S.n_steps[g];
        if (idx < forced_stride) a = forced[(size_t)g * forced_stride + idx];
      } else {
        unsigned long long key = S.best[g];
        if (key != 0ull) a = (int)(0xffffffffu - (uint32_t)(key & 0xffffffffull));
      }
      if (a >= N) a = -1;
    }
    s_action = a;
    S.best[g] = 0ull;
  }
  // init == 1: MvcEnv.s0;  init == 2: arbitrary state given as a removed-node mask (passed through `forced`) with the
  // caller's numCoveredEdges already in S.cov_edges (Predict on replay states, MultiDismantler_torch.py:263-302)
  const unsigned char* mask = reinterpret_cast<const unsigned char*>(forced);
  int n_masked = 0;
  for (int v = tid; v < St; v += ENV_T) {
    int lv;
    if (init == 1) lv = v < N ? 0 : -1;
    else if (init == 2) {
      const bool rm = v < N && mask[base + v];
      n_masked += rm;
      lv = (v < N && !rm) ? 0 : -1;
    } else lv = S.label[base + v] < 0 ? -1 : S.label[base + v] - base;
    lab[v] = lv;
  }
  if (init == 2) n_masked = block_sum_i(n_masked, red);
  __syncthreads();
  const int a = s_action;
  if (!init && (a < 0 || lab[a] < 0)) return;  // nothing to remove (uniform exit)

  ENV_TICK();  // 0: labels in
  // ---- stage the graph's CSR in shared memory ---------------------------------------------------------
  // A step starts from the ALIVE CSR the previous step emitted (rows = previous active nodes, entries = edges that were
  // still alive): dead edges never come back, so staging, hooking and compaction cost shrink with the graph.  s0 /
  // set_state start from the static CSR (rows = all nodes).  Either way row p stands for node arow[p].
  const bool from_alive = init == 0;
  const int nprev = from_alive ? S.n_act[g] : N;
  for (int p = tid; p < nprev; p += ENV_T) arow[p] = from_alive ? S.act_list[base + p] - base : p;
  int lay_pos[MD_MAX_LAYERS + 1];
  {
    int pos = 0;
    for (int ll = 0; ll < L; ++ll) {
      const int32_t* __restrict__ rp = (from_alive ? S.aptr[ll] : G.rowptr[ll]) + base;
      const int32_t* __restrict__ cl = from_alive ? S.acol[ll] : G.col[ll];
      const int32_t* __restrict__ sr = from_alive ? S.asrc[ll] : G.src[ll];
      const int e0 = rp[0], cnt = rp[nprev] - e0;
      for (int p = tid; p <= nprev; p += ENV_T) c_rp[ll * (St + 1) + p] = rp[p] - e0 + pos;
#pragma unroll 4
      for (int e = tid; e < cnt; e += ENV_T) {  // unrolled: the independent loads of four trips are in flight together
        c_cl[pos + e] = (IdxT)(cl[e0 + e] - base);
        c_src[pos + e] = (IdxT)(sr[e0 + e] - base);
      }
      lay_pos[ll] = pos;
      pos += cnt;
    }
    lay_pos[L] = pos;
  }

  if (!init) {
    // numCoveredEdges[l] += #neighbours still alive and not cut from a  (mvc_env.py:81-84) == deg_alive_l(a)
    if (tid < L) S.cov_edges[tid * G.n_graphs + g] += S.deg[tid][base + a];
    if (tid == 0) {
      s_lab_a = lab[a];
      lab[a] = -1;
    }
  }
  __syncthreads();

  ENV_TICK();  // 1: CSR staged
  // ---- Mcc.MCC: refine the partition layer by layer until every layer is stable ----------------------
  int stable = 0, l = 0;
  int rounds_dbg = 0;
  long long hook_dbg = 0, flat_dbg = 0, tq = 0;
  const bool incremental = init == 0 && inc_layout;  // opt-in: measured 2x SLOWER than the rounds at N = 1000 (below)
  if constexpr (INC) if (incremental) {
    // ---- incremental update (a step; s0 / set_state run the from-scratch rounds below) ---------------------------------
    // Removing a can only change the block that contained a, and inside it only what hangs on the edges that just died.
    // Invariant per layer l: every layer-l connected piece of a block contains a PENDING SOURCE of layer l, or the block is
    // layer-l connected.  Initially the pending sources are a's alive neighbours.  One task = (block x, layer l) with >= 2
    // pending sources: a multi-source search along the alive layer-l edges inside x, groups merging when they touch; it
    // stops as soon as at most one group still has a frontier.  Groups that ran out of frontier are complete pieces: they
    // leave x under fresh labels; the rest of x -- the group still searching plus everything never visited -- keeps x
    // untouched.  Every edge of the OTHER layers between a piece and the rest dies, and its end points become pending
    // sources there.  When nothing is pending every block is connected in every layer: the same fixpoint as the
    // reference's alternating full recomputation (Mcc.py:30-38), reached in work proportional to what changed.
    // (scripts/proto_incremental_lmcc.py checks the scheme against the from-scratch oracle; the GPU suite passes with it.)
    // MEASURED (C2, 256 x N = 1000): bit-exact, but the rollout takes 38.7 instead of 21.8 ms.  A step needs 5-15 search tasks
    // of ~4 levels each (cascades: every small piece that leaves the giant block re-opens the giant block in the other
    // layer), i.e. some hundred block barriers, and together the searches visit 1.5-3 x n nodes -- no less than the 3-4 edge
    // sweeps of the from-scratch rounds on graphs this small.  Kept behind md_debug_set_flags bit 16 as the reference
    // implementation of the scheme for graphs where n >> what a removal touches.
    int* own = tmp;   // [St] 0 = unvisited, else 1 + index of the source whose group claimed the node
    int* gpar = par;  // [St] union-find over the sources of the running task
    const int x0 = s_lab_a;
    for (int v = tid; v < St; v += ENV_T) {
      own[v] = 0;
      pend[v] = 0;
      rowof[v] = -1;
    }
    if (tid == 0) s_cnt[4] = N;  // fresh labels N, N + 1, ... (canonical labels are node ids < N)
    __syncthreads();
    for (int p = tid; p < nprev; p += ENV_T) rowof[arow[p]] = p;
    __syncthreads();
    {
      const int pa = rowof[a];
      if (pa >= 0)
        for (int ll = 0; ll < L; ++ll) {
          const int* rp = c_rp + ll * (St + 1);
          const int e1 = rp[pa + 1];
          for (int e = rp[pa] + tid; e < e1; e += ENV_T) {
            const int u = c_cl[e];
            if (lab[u] == x0) atomicOr(&pend[u], 1 << ll);
          }
        }
    }
    __syncthreads();
    int idle = 0;
    while (idle < L) {
      if (tid == 0) s_cnt[0] = 0;
      __syncthreads();
      for (int v = tid; v < N; v += ENV_T) {
        const int pb = pend[v];
        if (pb & (1 << l)) {
          pend[v] = pb & ~(1 << l);
          if (lab[v] >= 0) srcl[atomicAdd(&s_cnt[0], 1)] = v;
        }
      }
      __syncthreads();
      int rem = s_cnt[0];
      __syncthreads();
      if (rem == 0) {
        ++idle;
        l = (l + 1 == L) ? 0 : l + 1;
        continue;
      }
      idle = 0;
      const int* rp = c_rp + l * (St + 1);
      while (rem > 0) {  // one block at a time: the block of the first remaining source
        const int x = lab[srcl[0]];
        if (tid == 0) s_cnt[1] = 0, s_cnt[6] = 0;
        __syncthreads();
        for (int i = tid; i < rem; i += ENV_T) {
          const int v = srcl[i];
          if (lab[v] == x) vis[atomicAdd(&s_cnt[1], 1)] = v;
          else gnew[atomicAdd(&s_cnt[6], 1)] = v;
        }
        __syncthreads();
        const int ns = s_cnt[1];
        rem = s_cnt[6];
        for (int i = tid; i < rem; i += ENV_T) srcl[i] = gnew[i];
        __syncthreads();
        if (ns < 2) continue;  // one source: by the invariant its block is connected in this layer
        ++rounds_dbg;
        for (int i = tid; i < ns; i += ENV_T) {
          own[vis[i]] = i + 1;
          gpar[i] = i;
        }
        __syncthreads();
        int f0 = 0, f1 = ns, nroots = ns, nact = ns;
        while (nroots > 1 && nact > 1) {
          // expand the frontier vis[f0, f1): 8 lanes per node (hub rows do not serialise one thread)
          for (int i = f0 + (tid >> 3); i < f1; i += ENV_T / 8) {
            const int u = vis[i], gu = own[u] - 1, pu = rowof[u];
            const int e1 = rp[pu + 1];
            for (int e = rp[pu] + (tid & 7); e < e1; e += 8) {
              const int v = c_cl[e];
              if (lab[v] != x) continue;
              const int o = atomicCAS(&own[v], 0, gu + 1);
              if (o == 0) vis[atomicAdd(&s_cnt[1], 1)] = v;
              else if (o - 1 != gu) uf_unite(gpar, o - 1, gu);
            }
          }
          __syncthreads();
          f0 = f1;
          f1 = s_cnt[1];
          for (int i = tid; i < ns; i += ENV_T) gact[i] = 0;
          if (tid == 0) s_cnt[2] = 0, s_cnt[3] = 0;
          __syncthreads();
          for (int i = f0 + tid; i < f1; i += ENV_T) gact[uf_find(gpar, own[vis[i]] - 1)] = 1;
          __syncthreads();
          for (int i = tid; i < ns; i += ENV_T)
            if (uf_find(gpar, i) == i) {
              atomicAdd(&s_cnt[2], 1);
              if (gact[i]) atomicAdd(&s_cnt[3], 1);
            }
          __syncthreads();
          nroots = s_cnt[2];
          nact = s_cnt[3];
        }
        const int nvis = f1;
        if (nroots > 1) {
          // the block splits: the group that is still searching (or, if all ran dry, the group with the smallest node)
          // stays x, every other group is a complete piece and gets a fresh label
          if (tid == 0) s_cnt[5] = -1, s_cnt[7] = 0x7fffffff;
          for (int i = tid; i < ns; i += ENV_T) grmin[i] = 0x7fffffff;
          __syncthreads();
          if (nact == 1) {
            for (int i = tid; i < ns; i += ENV_T)
              if (gact[i] && uf_find(gpar, i) == i) s_cnt[5] = i;
          } else {
            for (int i = tid; i < nvis; i += ENV_T) atomicMin(&grmin[uf_find(gpar, own[vis[i]] - 1)], vis[i]);
            __syncthreads();
            for (int i = tid; i < ns; i += ENV_T)
              if (uf_find(gpar, i) == i) atomicMin(&s_cnt[7], grmin[i]);
            __syncthreads();
            for (int i = tid; i < ns; i += ENV_T)
              if (uf_find(gpar, i) == i && grmin[i] == s_cnt[7]) s_cnt[5] = i;
          }
          __syncthreads();
          const int keep = s_cnt[5];
          for (int i = tid; i < ns; i += ENV_T)
            if (i != keep && uf_find(gpar, i) == i) gnew[i] = atomicAdd(&s_cnt[4], 1);
          __syncthreads();
          for (int i = tid; i < nvis; i += ENV_T) {
            const int v = vis[i], r = uf_find(gpar, own[v] - 1);
            if (r != keep) lab[v] = gnew[r];
          }
          __syncthreads();
          // edges of the other layers that were alive inside x and cross labels now: both end points become pending
          for (int i = tid >> 3; i < nvis; i += ENV_T / 8) {
            const int v = vis[i], lv = lab[v];
            if (lv == x) continue;  // only nodes that moved
            const int pv = rowof[v];
            for (int l2 = 0; l2 < L; ++l2) {
              if (l2 == l) continue;
              const int* rp2 = c_rp + l2 * (St + 1);
              const int e1 = rp2[pv + 1];
              for (int e = rp2[pv] + (tid & 7); e < e1; e += 8) {
                const int u = c_cl[e], lu = lab[u];
                if (lu >= 0 && lu != lv && (lu == x || own[u] != 0)) {
                  atomicOr(&pend[v], 1 << l2);
                  atomicOr(&pend[u], 1 << l2);
                }
              }
            }
          }
          __syncthreads();
        }
        for (int i = tid; i < nvis; i += ENV_T) own[vis[i]] = 0;
        __syncthreads();
      }
      l = (l + 1 == L) ? 0 : l + 1;
    }
  }
  while (!incremental && stable < L) {
    for (int v = tid; v < N; v += ENV_T) par[v] = v;
    __syncthreads();
    tq = ENV_CLK();
    for (int e = lay_pos[l] + tid; e < lay_pos[l + 1]; e += ENV_T) {  // edge-parallel hooking
      const int v = c_src[e], u = c_cl[e];
      if (u < v) {
        const int lv = lab[v];
        // equal parents = same tree already: most entries of a sweep once the giant block has formed (two loads instead of
        // two root walks with their path-halving stores)
        if (lv >= 0 && lab[u] == lv && par[u] != par[v]) uf_unite(par, u, v);
      }
    }
    __syncthreads();
    hook_dbg += ENV_CLK() - tq;
    tq = ENV_CLK();
    int changed = 0;
    for (int v = tid; v < N; v += ENV_T) {
      int r = -1;
      if (lab[v] >= 0) {
        r = uf_find(par, v);
        changed |= (r != lab[v]);
      }
      tmp[v] = r;
    }
    changed = __syncthreads_or(changed);  // (also orders this round's label reads before the next round's writes)
    {
      int* t = lab;  // the new labels become `lab` without a copy pass (entries >= N of both arrays hold -1 / are never read)
      lab = tmp;
      tmp = t;
    }
    flat_dbg += ENV_CLK() - tq;
    stable = changed ? 1 : stable + 1;
    l = (l + 1 == L) ? 0 : l + 1;
    ++rounds_dbg;
  }
  ENV_TICK();  // 2: LMCC rounds
  ENV_PROF(if (blockIdx.x == 0 && tid == 0) g_env_dbg[16] = rounds_dbg, g_env_dbg[17] = hook_dbg, g_env_dbg[18] = flat_dbg;)
  (void)hook_dbg, (void)flat_dbg, (void)tq, (void)rounds_dbg;

  // ---- component sizes -> LMCC rank (Mcc.find_max_set_length) ------------------------------------------
  int rank;
  if constexpr (INC) {
    // Labels are node ids < N (from-scratch rounds) or fresh ids in [N, 2N) (incremental splits): one table pass counts the
    // members of every label and finds its smallest member, which becomes the canonical label again.
    int* ccnt = vis;   // [2 St] (vis + pend)
    int* cmin = srcl;  // [2 St] (srcl + gact)
    for (int i = tid; i < 2 * St; i += ENV_T) {
      ccnt[i] = 0;
      cmin[i] = 0x7fffffff;
    }
    __syncthreads();
    for (int v = tid; v < N; v += ENV_T)
      if (lab[v] >= 0) {
        atomicAdd(&ccnt[lab[v]], 1);
        atomicMin(&cmin[lab[v]], v);
      }
    __syncthreads();
    int mx = 0;
    for (int i = tid; i < 2 * St; i += ENV_T) mx = max(mx, ccnt[i]);
    rank = block_max_i(mx, red);
    for (int v = tid; v < N; v += ENV_T)
      if (lab[v] >= 0) lab[v] = cmin[lab[v]];
    __syncthreads();
  } else {
    for (int v = tid; v < N; v += ENV_T) tmp[v] = 0;
    __syncthreads();
    for (int v = tid; v < N; v += ENV_T)
      if (lab[v] >= 0) atomicAdd(&tmp[lab[v]], 1);
    __syncthreads();
    int mx = 0;
    for (int v = tid; v < N; v += ENV_T) mx = max(mx, tmp[v]);
    rank = block_max_i(mx, red);
  }

  ENV_TICK();  // 3: component sizes
  // ---- per layer: alive CSR, alive degree, status counters (get_status_info) -------------------------------
  // The alive CSR is a stream compaction of the staged CSR entries by the flag "both end labels equal": one
  // edge-parallel block scan per layer (hub rows cost nothing extra).  The running prefix is kept per entry (in the
  // c_src slots, dead after the prune), so deg(v) = pref[row end] - pref[row start] needs no second edge sweep.
  bool terminal = false;
  int n_act = 0, max_tot = 0;
  for (int ll = 0; ll < L; ++ll) {
    const int* rp = c_rp + ll * (St + 1);
    const int p0 = lay_pos[ll], cnt = lay_pos[ll + 1] - p0;
    const int e0 = G.rowptr[ll][base];  // the graph's region of acol[ll] / asrc[ll] is its static CSR region
    int32_t* __restrict__ aptr = S.aptr[ll] + base;
    int32_t* __restrict__ acol = S.acol[ll];
    int32_t* __restrict__ asrc = S.asrc[ll];
    IdxT* pref = c_src + p0;  // pref[i] = number of alive entries before entry i of this layer, i in [0, cnt]
    // thread t owns the contiguous entries [i0, i1): count, one block-wide exclusive scan of the 512 counts, then write.
    // (Reads and the in-place prefix writes of a thread stay inside its own range, so the two passes need no barrier
    // between them besides the scan's; the chunk-at-a-time version spent 2 barriers per 512 entries: 28 k cycles / step.)
    const int per = (cnt + ENV_T - 1) / ENV_T;
    const int i0 = min(cnt, tid * per), i1 = min(cnt, i0 + per);
    int mine = 0;
    for (int i = i0; i < i1; ++i) {
      const int lv = lab[c_src[p0 + i]];
      mine += (lv >= 0 && lab[c_cl[p0 + i]] == lv) ? 1 : 0;
    }
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += y;
    }
    __syncthreads();  // s_scan is free (previous layer's readers are done)
    if (lane == 31) s_scan[wid] = incl;
    __syncthreads();
    int woff = 0, running = 0;
#pragma unroll
    for (int k = 0; k < ENV_W; ++k) {
      const int c = s_scan[k];
      if (k < wid) woff += c;
      running += c;
    }
    {
      int pos = woff + incl - mine;
      for (int i = i0; i < i1; ++i) {
        const int u = c_cl[p0 + i], sv = c_src[p0 + i];
        const int lv = lab[sv];
        const bool flag = lv >= 0 && lab[u] == lv;
        pref[i] = (IdxT)pos;
        if (flag) {
          acol[e0 + pos] = u + base;
          asrc[e0 + pos] = sv + base;
          ++pos;
        }
      }
    }
    __syncthreads();
    const int tot = running;  // alive directed entries of the layer (= pref[cnt], which has no slot of its own)
    const int pend = p0 + cnt;
#define PREF_AT(x) ((x) == pend ? tot : (int)c_src[(x)])
    int dmax = 0;
    long long th = 0;
    for (int p = tid; p < nprev; p += ENV_T) {  // nodes outside the input rows keep their degree 0
      const int d = PREF_AT(rp[p + 1]) - PREF_AT(rp[p]);  // rp[] offsets are absolute positions in c_cl / c_src
      S.deg[ll][base + arow[p]] = d;
      par[p] = d;
      dmax = max(dmax, d);
      th += (long long)d * (d - 1) / 2;
    }
    const int dm = block_max_i(dmax, red);
    const long long t2 = block_sum_ll(th, red);
    if (tid == 0) {
      S.alive_edges[ll * G.n_graphs + g] = tot / 2;
      S.max_deg[ll * G.n_graphs + g] = dm;
      S.twohop[ll * G.n_graphs + g] = t2;
    }
    terminal |= (tot == 0);
    max_tot = max(max_tot, tot);
    __syncthreads();

    if (ll == 0) {
      // write labels back; compact the rows that are still active (ascending id, PrepareBatchGraph.py:128-138): tmp[q] =
      // input row of new row q
      for (int v = tid; v < St; v += ENV_T) S.label[base + v] = (v < N && lab[v] >= 0) ? lab[v] + base : -1;
      int run2 = 0;
      for (int p0r = 0; p0r < nprev; p0r += ENV_T) {
        const int p = p0r + tid;
        const int flag = (p < nprev && par[p] > 0) ? 1 : 0;
        const unsigned bal = __ballot_sync(0xffffffffu, flag);
        if (lane == 0) s_scan[wid] = __popc(bal);
        __syncthreads();
        int woff = 0, total = 0;
#pragma unroll
        for (int k = 0; k < ENV_W; ++k) {
          const int c = s_scan[k];
          if (k < wid) woff += c;
          total += c;
        }
        if (flag) {
          const int q = run2 + woff + __popc(bal & ((1u << lane) - 1u));
          S.act_list[base + q] = base + arow[p];
          tmp[q] = p;
        }
        run2 += total;
        __syncthreads();
      }
      n_act = run2;
    }
    // row pointers of the alive CSR in compact order
    for (int p = tid; p < n_act; p += ENV_T) aptr[p] = e0 + PREF_AT(rp[tmp[p]]);
    if (tid == 0) aptr[n_act] = e0 + tot;
#undef PREF_AT
    __syncthreads();
  }

  ENV_TICK();  // 4: alive CSR + degrees + compaction
#ifdef MD_PROFILE
  if (tid == 0 && g < 1024) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    g_env_cta[g][0] = (unsigned)(clock64() - t_cta0), g_env_cta[g][1] = rounds_dbg, g_env_cta[g][2] = lay_pos[L], g_env_cta[g][3] = smid;
  }
#endif
  // ---- bookkeeping: reward / AUDC in f64 with the reference's operation order ---------------------------
  if (tid == 0) {
    const int B = G.n_graphs;
    S.n_act[g] = n_act;
    S.rank[g] = rank;
    int nrem = init == 1 ? 0 : (init == 2 ? n_masked : S.n_removed[g] + 1);
    S.n_removed[g] = nrem;
    if (init) {
      S.score[g] = 0.0;
      S.n_steps[g] = 0;
      if (init == 1)
        for (int ll = 0; ll < L; ++ll) S.cov_edges[ll * B + g] = 0;
    } else {
      const double rk = (double)rank;
      const double mr = (double)G.max_rank[g];
      double r, mcc;
      if (S.cost_mode == 0) {
        // U/mvc_env.py:133-137: r_t = -rank / (max_rank * N); score += -r_t; MaxCCList.append(-r_t * N)
        r = __ddiv_rn(rk, __dmul_rn(mr, (double)N));
        mcc = __dmul_rn(r, (double)N);
      } else {
        // D/mvc_env.py:127-133: r_t = -rank / max_rank * (w0[a]/W0 + w1[a]/W1) / 2; MaxCCList.append(rank / max_rank)
        mcc = __ddiv_rn(rk, mr);
        r = __ddiv_rn(__dmul_rn(mcc, G.node_cost[base + a]), 2.0);
      }
      S.score[g] = __dadd_rn(S.score[g], r);
      const int t = S.n_steps[g];
      if (t < S.max_steps) {
        S.trace_action[(size_t)g * S.max_steps + t] = a;
        S.trace_rank[(size_t)g * S.max_steps + t] = rank;
        S.trace_maxcc[(size_t)g * S.max_steps + t] = mcc;
      }
      S.n_steps[g] = t + 1;
    }
    // aux features for the next forward (PrepareBatchGraph.py:92-100): quotients in f64, then cast to f32
    for (int ll = 0; ll < L; ++ll) {
      float* ax = S.aux + ((size_t)g * L + ll) * MD_AUX;
      const int E = G.n_edges[ll * B + g];
      ax[0] = (float)__ddiv_rn((double)nrem, (double)N);
      ax[1] = E > 0 ? (float)__ddiv_rn((double)S.cov_edges[ll * B + g], (double)E) : 0.f;
      ax[2] = (float)__ddiv_rn((double)S.twohop[ll * B + g], (double)((long long)N * (long long)N));
      ax[3] = 1.0f;
    }
    S.done[g] = terminal ? 1 : 0;
    if (terminal) {
      atomicAdd(S.n_done, 1);
      atomicAdd(&S.flags[FLAG_LEFT], 1);
    } else if (S.tail_n != nullptr && init == 0 && L == 2 && n_act <= TL_N && max_tot <= TL_E) {
      // small enough to finish on one SM: hand the graph over to md_rollout_tail.  For the lock-step kernels it is
      // finished from here on (they skip graphs with n_act == 0 / done != 0); n_done counts it when the tail completes it.
      S.tail_n[g] = n_act;
      S.n_act[g] = 0;
      S.done[g] = 1;
      atomicAdd(&S.flags[FLAG_LEFT], 1);
    }
  }
}

}  // namespace md

int md_env_step_large(const md_graphs_t* g, const md_state_t* s, const int32_t* forced, int32_t forced_stride,
                      int32_t forced_index, int init, md_stream_t stream);
extern int g_md_debug_flags;  // bit 16 here: incremental mutual-component update instead of the from-scratch prune rounds

extern "C" int md_debug_read_env_ctas(unsigned* host4096) {
  return (int)cudaMemcpyFromSymbol(host4096, md::g_env_cta, sizeof(unsigned) * 4096);
}
extern "C" int md_debug_read_env(long long* host, int n) {
  return (int)cudaMemcpyFromSymbol(host, md::g_env_dbg, sizeof(long long) * n);
}
static int g_force_large = 0;
extern "C" int32_t md_env_small_max_stride(void) { return g_force_large ? 0 : md::ENV_SMALL_MAX; }
// 1 if this batch takes the single-CTA (shared-memory) env kernel, 0 for the cooperative global-memory kernel
static bool env_idx16(const md_graphs_t* g) { return g->max_graph_nnz < 65536 && g->stride < 65536; }
static size_t env_small_smem(const md_graphs_t* g) {
  const size_t idx_bytes = env_idx16(g) ? 2 : 4;
  const size_t arrays = (g_md_debug_flags & 65536) ? 11 : 4;  // + 7 arrays of the incremental update when it is switched on
  return (arrays * g->stride + (size_t)g->n_layers * (g->stride + 1)) * sizeof(int) +
         ((2 * (size_t)g->max_graph_nnz * idx_bytes + 15) & ~(size_t)15);
}
extern "C" int32_t md_env_uses_small_path(const md_graphs_t* g) {
  return !(g_force_large || g->stride > md::ENV_SMALL_MAX || g->max_graph_nnz <= 0 || env_small_smem(g) > md::ENV_SMALL_SMEM);
}
extern "C" void md_debug_force_large_env(int32_t on) { g_force_large = on; }

static int env_launch(const md_graphs_t* g, const md_state_t* s, const int32_t* forced, int32_t forced_stride,
                      int32_t forced_index, int init, md_stream_t stream) {
  if (!g || !s) return MD_EINVAL;
  if (g->n_layers < 1 || g->n_layers > MD_MAX_LAYERS) return MD_ELAYERS;
  if (g->stride % MD_TILE_M != 0 || g->n_total != g->n_graphs * g->stride) return MD_EINVAL;
  const size_t smem = env_small_smem(g);
  if (!md_env_uses_small_path(g)) return md_env_step_large(g, s, forced, forced_stride, forced_index, init, stream);
  const bool i16 = env_idx16(g);
  const bool inc = (g_md_debug_flags & 65536) != 0;
  const void* fn = i16 ? (inc ? (const void*)md::env_step_small<uint16_t, true> : (const void*)md::env_step_small<uint16_t, false>)
                       : (inc ? (const void*)md::env_step_small<int32_t, true> : (const void*)md::env_step_small<int32_t, false>);
  static md::PerDevice<size_t> configured_pd[4];
  size_t& configured_smem = configured_pd[2 * i16 + inc].get();
  if (smem > 48 * 1024 && smem > configured_smem) {
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    configured_smem = smem;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (init) {
    cudaError_t e = cudaMemsetAsync(s->n_done, 0, sizeof(int32_t), st);
    if (e != cudaSuccess) return (int)e;
    if (s->flags) {
      e = cudaMemsetAsync(s->flags + md::FLAG_LEFT, 0, sizeof(int32_t), st);
      if (e != cudaSuccess) return (int)e;
    }
  }
  md_graphs_t gg = *g;
  md_state_t ss = *s;
  int max_nnz = g->max_graph_nnz;
  void* args[] = {&gg, &ss, (void*)&forced, &forced_stride, &forced_index, &init, &max_nnz};
  cudaError_t e = cudaLaunchKernel(fn, dim3(g->n_graphs), dim3(md::ENV_T), args, smem, st);
  if (e != cudaSuccess) return (int)e;
  MD_CHECK_LAUNCH();
  return 0;
}

extern "C" int md_env_reset(const md_graphs_t* g, const md_state_t* s, md_stream_t stream) {
  return env_launch(g, s, nullptr, 0, -1, 1, stream);
}

extern "C" int md_env_set_state(const md_graphs_t* g, const md_state_t* s, const uint8_t* removed_mask,
                                md_stream_t stream) {
  if (!removed_mask) return MD_EINVAL;
  return env_launch(g, s, reinterpret_cast<const int32_t*>(removed_mask), 0, -1, 2, stream);
}

extern "C" int md_env_step(const md_graphs_t* g, const md_state_t* s, const int32_t* actions, int32_t action_stride,
                           int32_t action_index, md_stream_t stream) {
  return env_launch(g, s, actions, action_stride, action_index, 0, stream);
}
