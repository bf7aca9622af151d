// Environment step for graphs too large for one CTA's shared memory (e.g. config C4: N = 1M, 3 layers):
// one persistent cooperative kernel, all state in global memory (labels / parents stay L2-resident: 4 MB at N=1M),
// grid-wide barriers between the phases, no host round trip inside a step.
//
// Same semantics and reference citations as md_env.cu (mvc_env.py:31-52,74-87,133-162; Mcc.py:3-38;
// PrepareBatchGraph.py:35-100).  LMCC = iterative mutual-component pruning: per layer an ECL-CC style union-find
// (hook the larger root under the smaller with atomicCAS, path halving) restricted to edges whose end points
// carry the same label, repeated round-robin over the layers until L consecutive layers leave the labels unchanged.
#include "md_common.cuh"

namespace md {

constexpr int LG_T = 256;

__device__ __forceinline__ int ld_volatile(const int* p) { return *reinterpret_cast<const volatile int*>(p); }

// Parent reads are plain (L1-cached) loads on purpose: in a graph with a giant component every find ends at the same
// root word, and a million volatile reads of one word serialise in one L2 slice (measured: 0.6 ms per hooking sweep at
// N = 1M, whatever the number of edges hooked).  A stale line only shows an OLDER parent, which is still an ancestor
// (parents only ever move towards smaller ids), so find still climbs towards the root; a stale root is caught by the
// atomicCAS in guf_unite, whose return value is always current.
__device__ __forceinline__ int guf_find(int* par, int x) {
  int p = par[x];
  while (p != x) {
    const int gp = par[p];
    if (gp != p) par[x] = gp;  // path halving (benign race)
    x = p;
    p = gp;
  }
  return x;
}

__device__ __forceinline__ void guf_unite(int* par, int a, int b) {
  int ra = guf_find(par, a), rb = guf_find(par, b);
  while (ra != rb) {
    const int hi = max(ra, rb), lo = min(ra, rb);
    const int old = atomicCAS(par + hi, hi, lo);
    if (old == hi) break;
    ra = guf_find(par, old);
    rb = guf_find(par, lo);
  }
}

// sense-free monotone grid barrier; all CTAs are co-resident (cooperative launch)
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned& epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    epoch += 1;
    const unsigned target = epoch * gridDim.x;
    __threadfence();
    atomicAdd(counter, 1u);
    while (*reinterpret_cast<volatile unsigned*>(counter) < target) {
    }
    __threadfence();
  }
  __syncthreads();
}

// flags slots of the incremental update: counters [9..15] (sources / rest of the round (two parities), visited, roots,
// active roots, kept root, smallest node), [16] next fresh label (persists from launch to launch), [17..20] pending
// sources per layer
constexpr int LF_SRC = 9, LF_VIS = 10, LF_ROOTS = 11, LF_ACT = 12, LF_KEEP = 13, LF_MIN = 14, LF_REST = 15, LF_NEXT = 16,
              LF_PEND = 17, LF_SRC2 = 21, LF_ROOTS2 = 22, LF_ACT2 = 23;

__global__ void __launch_bounds__(LG_T) env_step_large(md_graphs_t G, md_state_t S, const int32_t* forced,
                                                      int forced_stride, int forced_index, int init, int cpg,
                                                      int chunk, int incremental) {
  const int tid = threadIdx.x;
  const int gtid = blockIdx.x * LG_T + tid;
  const int gsz = gridDim.x * LG_T;
  const int St = G.stride, B = G.n_graphs, L = G.n_layers, NT = G.n_total;
  unsigned* bar = reinterpret_cast<unsigned*>(S.flags);  // flags[0] barrier counter, [1..3] changed flags
  int* flags = S.flags;
  unsigned epoch = 0;
  int* par = S.parent;
  int* label = S.label;
  __shared__ int s_red[LG_T / 32];
  __shared__ int s_off;

  // ---- P0: pick / validate the action, apply it --------------------------------------------------------
  const unsigned char* mask = reinterpret_cast<const unsigned char*>(forced);  // init == 2: removed-node mask
  if (init) {
    for (int g = gtid; g < B; g += gsz) S.n_removed[g] = 0;
  }
  if (init == 2) grid_barrier(bar, epoch);
  if (init) {
    for (int v = gtid; v < NT; v += gsz) {
      const int g = v / St;
      const bool inside = (v - g * St) < G.n_nodes[g];
      const bool rm = init == 2 && inside && mask[v];
      label[v] = (inside && !rm) ? g * St : -1;
      if (rm) atomicAdd(&S.n_removed[g], 1);
    }
  }
  for (int g = gtid; g < B; g += gsz) {
    int a = -1;
    if (!init && !S.done[g]) {
      if (forced != nullptr) {
        const int idx = forced_index >= 0 ? forced_index : S.n_steps[g];
        if (idx < forced_stride) a = forced[(size_t)g * forced_stride + idx];
      } else {
        const unsigned long long key = S.best[g];
        if (key != 0ull) a = (int)(0xffffffffu - (uint32_t)(key & 0xffffffffull));
      }
      if (a >= G.n_nodes[g]) a = -1;
      if (a >= 0 && label[g * St + a] < 0) a = -1;
      if (a >= 0) {
        for (int l = 0; l < L; ++l) S.cov_edges[l * B + g] += S.deg[l][g * St + a];
        if (incremental) {  // pending sources of the incremental update: a's alive neighbours, layer by layer
          const int va = g * St + a, x0 = label[va];
          int* pbits = S.lg_scratch + NT;
          for (int l = 0; l < L; ++l) {
            const int e1 = G.rowptr[l][va + 1];
            for (int e = G.rowptr[l][va]; e < e1; ++e) {
              const int u = G.col[l][e];
              if (label[u] == x0 && !(atomicOr(&pbits[u], 1 << l) & (1 << l))) atomicAdd(&flags[LF_PEND + l], 1);
            }
          }
        }
        label[g * St + a] = -1;
      }
    }
    S.cur_action[g] = a;
    S.best[g] = 0ull;
    if (init && g == 0) {
      flags[LF_NEXT] = NT;  // fresh labels of the incremental update start above the node ids
      for (int l = 0; l < MD_MAX_LAYERS; ++l) flags[LF_PEND + l] = 0;
    }
    S.rank[g] = (init || a >= 0) ? 0 : S.rank[g];
    if (init || a >= 0)
      for (int l = 0; l < L; ++l) {
        S.alive_edges[l * B + g] = 0;
        S.max_deg[l * B + g] = 0;
        S.twohop[l * B + g] = 0;
      }
  }
  // flags[1..3]: "labels changed" flag of round r lives in slot r % 3 (memset to 0 by the host before the launch);
  // a slot is re-cleared two rounds after its use, when every CTA has certainly read it.
  grid_barrier(bar, epoch);

  // phase timers (block 0, thread 0): flags[6] = MCC rounds in ns, flags[7] = rounds << 24 | (sizes + degrees in us);
  // flags[4..5] belong to md_qhead's tile queue
  auto now_ns = []() -> unsigned long long {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
  };
  unsigned long long t_ph = now_ns();
  // ---- Mcc.MCC --------------------------------------------------------------------------------------------
  int stable = 0, l = 0, round = 0;
  if (incremental && !init) {
    // ---- incremental update (same scheme as md_env.cu, grid-wide): only what hangs on the edges that just died is visited.
    // At N = 1M a removal makes the searches visit some 10^4 nodes instead of the 3 x 4 sweeps over 10^6 nodes of the
    // from-scratch rounds below.  Labels: node ids (< NT) from the from-scratch rounds, fresh ids (>= NT, one counter per
    // rollout) for pieces that split off; the rest of a block keeps its label without being visited.
    int* own = S.lg_scratch;  // [NT] 0 = unvisited, else 1 + index of the claiming source   (all zero between launches)
    int* pbits = own + NT;    // [NT] pending-source bits, one per layer                      (all zero between launches)
    int* vis = pbits + NT;    // [NT] visited list; its tail segment is the current frontier
    int* srcl = vis + NT;     // [NT] pending sources of the running round
    int* tmpl = srcl + NT;    // [NT] partition scratch
    int* gpar = tmpl + NT;    // [NT] union-find over the sources of the running task
    int* gact = gpar + NT;    // [NT] per source group: has a frontier
    int* grmin = gact + NT;   // [NT] per source group: smallest node
    int* gnew = grmin + NT;   // [NT] per source group: fresh label
    int idle = 0, parity = 0;
    while (idle < L) {
      if (ld_volatile(&flags[LF_PEND + l]) == 0) {  // (last modified before a grid barrier: every CTA reads the same value)
        ++idle;
        l = (l + 1 == L) ? 0 : l + 1;
        continue;
      }
      idle = 0;
      const int* __restrict__ rp = G.rowptr[l];
      const int* __restrict__ cl = G.col[l];
      int* c_src = flags + (parity ? LF_SRC2 : LF_SRC);
      parity ^= 1;
      if (gtid == 0) *c_src = 0;
      grid_barrier(bar, epoch);
      if (gtid == 0) flags[LF_PEND + l] = 0;  // (every CTA has read it: nothing adds to this layer's count during its own round)
      for (int v = gtid; v < NT; v += gsz) {
        const int pb = pbits[v];
        if (pb & (1 << l)) {
          pbits[v] = pb & ~(1 << l);
          if (label[v] >= 0) srcl[atomicAdd(c_src, 1)] = v;
        }
      }
      grid_barrier(bar, epoch);
      int rem = ld_volatile(c_src);
      while (rem > 0) {  // one block at a time: the block of the first remaining source
        const int x = label[srcl[0]];
        if (gtid == 0) flags[LF_VIS] = 0, flags[LF_REST] = 0;
        grid_barrier(bar, epoch);
        for (int i = gtid; i < rem; i += gsz) {
          const int v = srcl[i];
          if (label[v] == x) vis[atomicAdd(&flags[LF_VIS], 1)] = v;
          else tmpl[atomicAdd(&flags[LF_REST], 1)] = v;
        }
        grid_barrier(bar, epoch);
        const int ns = ld_volatile(&flags[LF_VIS]);
        rem = ld_volatile(&flags[LF_REST]);
        for (int i = gtid; i < rem; i += gsz) srcl[i] = tmpl[i];
        if (ns >= 2)
          for (int i = gtid; i < ns; i += gsz) {
            own[vis[i]] = i + 1;
            gpar[i] = i;
          }
        grid_barrier(bar, epoch);
        if (ns < 2) continue;  // one source: by the invariant its block is connected in this layer
        ++round;
        int f0 = 0, f1 = ns, nroots = ns, nact = ns, lvl = 0;
        while (nroots > 1 && nact > 1) {
          // two grid barriers per level: (1) expand the frontier; (2) count the groups and those that still have a frontier.
          // The counters of a level are cleared during the expansion of the NEXT level (two parities: a CTA that is late
          // reading level k's counts cannot see level k + 1's reset of the same words).
          int* c_roots = flags + ((lvl & 1) ? LF_ROOTS2 : LF_ROOTS);
          int* c_act = flags + ((lvl & 1) ? LF_ACT2 : LF_ACT);
          if (gtid == 0) *c_roots = 0, *c_act = 0;
          for (int i = gtid; i < ns; i += gsz) gact[i] = 0;
          for (int i = f0 + (gtid >> 3); i < f1; i += gsz / 8) {  // 8 lanes per frontier node
            const int u = vis[i], gu = own[u] - 1;
            const int e1 = rp[u + 1];
            for (int e = rp[u] + (gtid & 7); e < e1; e += 8) {
              const int v = cl[e];
              if (label[v] != x) continue;
              const int o = atomicCAS(&own[v], 0, gu + 1);
              if (o == 0) vis[atomicAdd(&flags[LF_VIS], 1)] = v;
              else if (o - 1 != gu) guf_unite(gpar, o - 1, gu);
            }
          }
          grid_barrier(bar, epoch);
          f0 = f1;
          f1 = ld_volatile(&flags[LF_VIS]);
          for (int i = f0 + gtid; i < f1; i += gsz) {  // (no unions in this phase: roots are stable)
            const int r = guf_find(gpar, own[vis[i]] - 1);
            if (atomicExch(&gact[r], 1) == 0) atomicAdd(c_act, 1);
          }
          for (int i = gtid; i < ns; i += gsz)
            if (guf_find(gpar, i) == i) atomicAdd(c_roots, 1);
          grid_barrier(bar, epoch);
          nroots = ld_volatile(c_roots);
          nact = ld_volatile(c_act);
          ++lvl;
        }
        const int nvis = f1;
        if (nroots > 1) {
          // split: the group that is still searching (or, if all ran dry, the one with the smallest node) stays x
          if (gtid == 0) flags[LF_KEEP] = -1, flags[LF_MIN] = 0x7fffffff;
          for (int i = gtid; i < ns; i += gsz) grmin[i] = 0x7fffffff;
          grid_barrier(bar, epoch);
          if (nact == 1) {
            for (int i = gtid; i < ns; i += gsz)
              if (gact[i] && guf_find(gpar, i) == i) flags[LF_KEEP] = i;
          } else {
            for (int i = gtid; i < nvis; i += gsz) atomicMin(&grmin[guf_find(gpar, own[vis[i]] - 1)], vis[i]);
            grid_barrier(bar, epoch);
            for (int i = gtid; i < ns; i += gsz)
              if (guf_find(gpar, i) == i) atomicMin(&flags[LF_MIN], grmin[i]);
            grid_barrier(bar, epoch);
            const int mn = ld_volatile(&flags[LF_MIN]);
            for (int i = gtid; i < ns; i += gsz)
              if (guf_find(gpar, i) == i && grmin[i] == mn) flags[LF_KEEP] = i;
          }
          grid_barrier(bar, epoch);
          const int keep = ld_volatile(&flags[LF_KEEP]);
          for (int i = gtid; i < ns; i += gsz)
            if (i != keep && guf_find(gpar, i) == i) gnew[i] = atomicAdd(&flags[LF_NEXT], 1);
          grid_barrier(bar, epoch);
          for (int i = gtid; i < nvis; i += gsz) {
            const int v = vis[i], r = guf_find(gpar, own[v] - 1);
            if (r != keep) label[v] = gnew[r];
          }
          grid_barrier(bar, epoch);
          // edges of the other layers that were alive inside x and cross labels now: both end points become pending
          for (int i = gtid >> 3; i < nvis; i += gsz / 8) {
            const int v = vis[i], lv = label[v];
            if (lv == x) continue;  // only nodes that moved
            for (int l2 = 0; l2 < L; ++l2) {
              if (l2 == l) continue;
              const int* __restrict__ rp2 = G.rowptr[l2];
              const int* __restrict__ cl2 = G.col[l2];
              const int e1 = rp2[v + 1];
              for (int e = rp2[v] + (gtid & 7); e < e1; e += 8) {
                const int u = cl2[e], lu = label[u];
                if (lu >= 0 && lu != lv && (lu == x || own[u] != 0)) {
                  if (!(atomicOr(&pbits[v], 1 << l2) & (1 << l2))) atomicAdd(&flags[LF_PEND + l2], 1);
                  if (!(atomicOr(&pbits[u], 1 << l2) & (1 << l2))) atomicAdd(&flags[LF_PEND + l2], 1);
                }
              }
            }
          }
          grid_barrier(bar, epoch);
        }
        for (int i = gtid; i < nvis; i += gsz) own[vis[i]] = 0;
        grid_barrier(bar, epoch);
      }
      l = (l + 1 == L) ? 0 : l + 1;
    }
    stable = L;  // skip the from-scratch rounds
  }
  while (stable < L) {
    const int32_t* __restrict__ rp = G.rowptr[l];
    const int32_t* __restrict__ cl = G.col[l];
    int* chg = flags + 1 + (round % 3);
    // Afforest-style round (Sutton et al.): link every node to its first two same-label neighbours, compress, find the
    // dominant set of each graph by sampling, then finish only the rows OUTSIDE that set -- an edge between a skipped
    // row and an outside node is still seen from the outside node's row (the CSR is symmetric), and edges inside the
    // dominant set join nodes that are already together.  A full hooking sweep over the 8M entries of a layer cost 0.6 ms.
    int* smp = S.comp_size;  // sample counters per root (zeroed again before the component sizes are counted)
    for (int v = gtid; v < NT; v += gsz) {
      par[v] = v;
      smp[v] = 0;
    }
    for (int g = gtid; g < B; g += gsz) S.best[g] = 0ull;  // scratch: (samples << 32 | root) of the dominant set
    if (gtid == 0) flags[1 + ((round + 1) % 3)] = 0;  // slot of the next round (last used in round - 2)
    grid_barrier(bar, epoch);
    for (int v = gtid; v < NT; v += gsz) {
      const int lv = label[v];
      if (lv < 0) continue;
      const int e0 = rp[v], e1 = min(rp[v + 1], e0 + 2);
      for (int e = e0; e < e1; ++e) {
        const int u = cl[e];
        if (label[u] == lv) guf_unite(par, u, v);
      }
    }
    grid_barrier(bar, epoch);
    for (int v = gtid; v < NT; v += gsz) {
      int r = -1;
      if (label[v] >= 0) {
        r = guf_find(par, v);
        par[v] = r;
      }
      if (((v >> 5) & 15) == 0) {  // every 16th warp samples; one atomic per distinct root of the warp
        const unsigned same = __match_any_sync(0xffffffffu, r);
        if (r >= 0 && (int)(__ffs(same) - 1) == (tid & 31)) {
          const unsigned long long c = (unsigned long long)(atomicAdd(smp + r, __popc(same)) + __popc(same));
          atomicMax(reinterpret_cast<unsigned long long*>(S.best) + v / St, (c << 32) | (unsigned)r);
        }
      }
    }
    grid_barrier(bar, epoch);
    for (int v = gtid; v < NT; v += gsz) {
      const int lv = label[v];
      if (lv < 0) continue;
      const int dom = (int)(unsigned)(ld_volatile(reinterpret_cast<const int*>(S.best + v / St)));  // low word = root
      if (par[v] == dom) continue;  // (compressed in the previous sweep; membership never ends)
      const int e1 = rp[v + 1];
      for (int e = rp[v]; e < e1; ++e) {
        const int u = cl[e];
        if (label[u] == lv) guf_unite(par, u, v);
      }
    }
    grid_barrier(bar, epoch);
    int changed = 0;
    for (int v = gtid; v < NT; v += gsz) {
      const int lv = label[v];
      if (lv < 0) continue;
      const int r = guf_find(par, v);
      if (r != lv) {
        label[v] = r;  // only this thread reads/writes label[v] in this phase
        changed = 1;
      }
    }
    if (__syncthreads_or(changed) && tid == 0) atomicOr(chg, 1);
    grid_barrier(bar, epoch);
    const int any = ld_volatile(chg);
    stable = any ? 1 : stable + 1;
    l = (l + 1 == L) ? 0 : l + 1;
    ++round;
  }

  if (gtid == 0) {
    const unsigned long long t = now_ns();
    flags[6] = (int)(t - t_ph);
    t_ph = t;
  }
  // ---- component sizes, alive degrees, per-graph counters -------------------------------------------------
  int* csz = S.comp_size;
  for (int g = gtid; g < B; g += gsz) S.best[g] = 0ull;  // (scratch of the MCC rounds; the arg-max key starts from 0)
  const int n_labels = incremental ? 3 * NT : NT;  // incremental update: fresh labels in [NT, 3 NT)
  for (int v = gtid; v < n_labels; v += gsz) csz[v] = 0;
  grid_barrier(bar, epoch);
  for (int v = gtid; v < NT; v += gsz) {
    const int lv = label[v];
    {  // one atomic per distinct label of the warp (the giant component would otherwise serialise ~N atomics on one word)
      const unsigned same = __match_any_sync(__activemask(), lv);
      if (lv >= 0 && (int)(__ffs(same) - 1) == (tid & 31)) atomicAdd(csz + lv, __popc(same));
    }
    const int g = v / St;
    if (!init && S.cur_action[g] < 0) continue;  // untouched graph: keep its counters
    for (int ll = 0; ll < L; ++ll) {
      int d = 0;
      if (lv >= 0) {
        const int32_t* __restrict__ rp = G.rowptr[ll];
        const int32_t* __restrict__ cl = G.col[ll];
        const int e1 = rp[v + 1];
        for (int e = rp[v]; e < e1; ++e) d += (label[cl[e]] == lv);
      }
      S.deg[ll][v] = d;
      // warp-aggregated when the whole warp sits in one graph (always true for large graphs)
      const int g0 = __shfl_sync(0xffffffffu, g, 0);
      if (__all_sync(0xffffffffu, g == g0)) {
        const int ds = warp_sum_i(d), dm = warp_max_i(d);
        const long long th = warp_sum_ll((long long)d * (d - 1) / 2);
        if ((tid & 31) == 0 && ds) {
          atomicAdd(&S.alive_edges[ll * B + g], ds);  // holds 2x the edge count until the epilogue halves it
          atomicMax(&S.max_deg[ll * B + g], dm);
          atomicAdd(reinterpret_cast<unsigned long long*>(&S.twohop[ll * B + g]), (unsigned long long)th);
        }
      } else if (d) {
        atomicAdd(&S.alive_edges[ll * B + g], d);
        atomicMax(&S.max_deg[ll * B + g], d);
        atomicAdd(reinterpret_cast<unsigned long long*>(&S.twohop[ll * B + g]), (unsigned long long)((long long)d * (d - 1) / 2));
      }
    }
  }
  grid_barrier(bar, epoch);
  for (int v = gtid; v < NT; v += gsz) {
    const int g = v / St;
    if (!init && S.cur_action[g] < 0) continue;
    const int lv2 = label[v];
    int c = lv2 >= 0 ? csz[lv2] : 0;  // every node looks up the size of its own block (labels need not be node ids)
    const int g0 = __shfl_sync(0xffffffffu, g, 0);
    if (__all_sync(0xffffffffu, g == g0)) {
      c = warp_max_i(c);
      if ((tid & 31) == 0 && c) atomicMax(&S.rank[g], c);
    } else if (c) {
      atomicMax(&S.rank[g], c);
    }
  }

  if (gtid == 0) {
    const unsigned long long t = now_ns();
    flags[7] = (round << 24) | (int)min((t - t_ph) / 1000ull, (unsigned long long)0xffffff);
  }
  // ---- compaction of the active nodes + alive CSR: per-chunk counts / degree sums, then ordered write ---------
  // chunk c covers nodes [g*St + k*chunk, g*St + min((k+1)*chunk, St)) with g = c / cpg, k = c % cpg
  int* ccount = par;  // parents are dead now; (1 + L) * B * cpg entries: [0] active count, [1 + l] alive-degree sum
  const int nchunks = B * cpg;
  for (int c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const int g = c / cpg, k = c - g * cpg;
    const int v0 = g * St + k * chunk, v1 = g * St + min((k + 1) * chunk, St);
    for (int q = 0; q <= L; ++q) {
      int cnt = 0;
      if (q == 0) {
        for (int v = v0 + tid; v < v1; v += LG_T) cnt += (S.deg[0][v] > 0);
      } else {
        for (int v = v0 + tid; v < v1; v += LG_T) cnt += S.deg[q - 1][v];
      }
      cnt = warp_sum_i(cnt);
      __syncthreads();
      if ((tid & 31) == 0) s_red[tid >> 5] = cnt;
      __syncthreads();
      if (tid == 0) {
        int t = 0;
        for (int w = 0; w < LG_T / 32; ++w) t += s_red[w];
        ccount[q * nchunks + c] = t;
      }
    }
    __syncthreads();
  }
  grid_barrier(bar, epoch);
  for (int c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const int g = c / cpg, k = c - g * cpg;
    if (!init && S.cur_action[g] < 0) continue;  // uniform per block
    const int v0 = g * St + k * chunk, v1 = g * St + min((k + 1) * chunk, St);
    const int lane = tid & 31, wid = tid >> 5;
    for (int q = 0; q <= L; ++q) {
      // offset of this chunk inside its graph for quantity q
      int part = 0;
      for (int j = tid; j < k; j += LG_T) part += ld_volatile(ccount + q * nchunks + g * cpg + j);
      part = warp_sum_i(part);
      __syncthreads();
      if (lane == 0) s_red[wid] = part;
      __syncthreads();
      if (tid == 0) {
        int t = 0;
        for (int w = 0; w < LG_T / 32; ++w) t += s_red[w];
        s_off = t;
      }
      __syncthreads();
      int running = s_off;
      __syncthreads();
      if (q == 0) {
        if (tid == 0 && k == cpg - 1) S.n_act[g] = running + ld_volatile(ccount + c);
        for (int vb = v0; vb < v1; vb += LG_T) {
          const int v = vb + tid;
          const int flag = (v < v1 && S.deg[0][v] > 0) ? 1 : 0;
          const unsigned bal = __ballot_sync(0xffffffffu, flag);
          __syncthreads();
          if (lane == 0) s_red[wid] = __popc(bal);
          __syncthreads();
          int woff = 0, total = 0;
#pragma unroll
          for (int w = 0; w < LG_T / 32; ++w) {
            const int cc = s_red[w];
            if (w < wid) woff += cc;
            total += cc;
          }
          const int p = running + woff + __popc(bal & ((1u << lane) - 1u));
          if (flag) {
            S.act_list[g * St + p] = v;
            csz[v] = p;  // component sizes are consumed; remember the compact position for the alive CSR
          }
          running += total;
        }
      } else {
        const int ll = q - 1;
        const int32_t* __restrict__ rp = G.rowptr[ll];
        const int32_t* __restrict__ cl = G.col[ll];
        const int e0 = rp[g * St];
        int32_t* __restrict__ aptr = S.aptr[ll] + g * St;
        int32_t* __restrict__ acol = S.acol[ll];
        if (tid == 0 && k == cpg - 1) {
          // n_act of this graph = offset of the last chunk + its count (recomputed: q == 0 pass already stored it)
          int nact = 0;
          for (int j = 0; j < cpg; ++j) nact += ld_volatile(ccount + g * cpg + j);
          aptr[nact] = e0 + running + ld_volatile(ccount + q * nchunks + c);
        }
        for (int vb = v0; vb < v1; vb += LG_T) {
          const int v = vb + tid;
          const int d = v < v1 ? S.deg[ll][v] : 0;
          int incl = d;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += n;
          }
          __syncthreads();
          if (lane == 31) s_red[wid] = incl;
          __syncthreads();
          int woff = 0, total = 0;
#pragma unroll
          for (int w = 0; w < LG_T / 32; ++w) {
            const int cc = s_red[w];
            if (w < wid) woff += cc;
            total += cc;
          }
          if (d > 0) {
            int o = e0 + running + woff + incl - d;
            aptr[csz[v]] = o;
            const int lv = label[v];
            const int e1 = rp[v + 1];
            for (int e = rp[v]; e < e1; ++e) {
              const int u = cl[e];
              if (label[u] == lv) acol[o++] = u;
            }
          }
          running += total;
        }
      }
      __syncthreads();
    }
  }
  grid_barrier(bar, epoch);

  // ---- per-graph bookkeeping (identical arithmetic to md_env.cu) ----------------------------------------------
  for (int g = gtid; g < B; g += gsz) {
    const int a = S.cur_action[g];
    if (!init && a < 0) continue;
    const int N = G.n_nodes[g];
    const int rank = S.rank[g];
    bool terminal = false;
    for (int ll = 0; ll < L; ++ll) {
      const int e2 = S.alive_edges[ll * B + g];
      S.alive_edges[ll * B + g] = e2 / 2;
      terminal |= (e2 == 0);
    }
    const int nrem = init ? S.n_removed[g] : S.n_removed[g] + 1;  // init: 0 (s0) or the mask population
    S.n_removed[g] = nrem;
    if (init) {
      S.score[g] = 0.0;
      S.n_steps[g] = 0;
      if (init == 1)
        for (int ll = 0; ll < L; ++ll) S.cov_edges[ll * B + g] = 0;
    } else {
      const double rk = (double)rank, mr = (double)G.max_rank[g];
      double r, mcc;
      if (S.cost_mode == 0) {
        r = __ddiv_rn(rk, __dmul_rn(mr, (double)N));
        mcc = __dmul_rn(r, (double)N);
      } else {
        mcc = __ddiv_rn(rk, mr);
        r = __ddiv_rn(__dmul_rn(mcc, G.node_cost[g * St + a]), 2.0);
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
    }
  }
}

}  // namespace md

extern int g_md_debug_flags;  // bit 17 here: from-scratch prune rounds on every step (A/B reference of the incremental update)

int md_env_step_large(const md_graphs_t* g, const md_state_t* s, const int32_t* forced, int32_t forced_stride,
                      int32_t forced_index, int init, md_stream_t stream) {
  if (!s->parent || !s->comp_size || !s->flags || !s->cur_action) return MD_EINVAL;
  static md::PerDevice<int> grid_pd;
  int& grid = grid_pd.get();
  if (grid == 0) {
    int dev = 0, nsm = 0, occ = 0, coop = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    if (!coop) return MD_EINVAL;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, md::env_step_large, md::LG_T, 0);
    if (e != cudaSuccess) return (int)e;
    if (occ > 4) occ = 4;
    grid = nsm * occ;
  }
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(s->flags, 0, 8 * sizeof(int32_t), st);
  if (e != cudaSuccess) return (int)e;
  if (init) {
    e = cudaMemsetAsync(s->n_done, 0, sizeof(int32_t), st);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemsetAsync(s->flags + md::FLAG_LEFT, 0, sizeof(int32_t), st);
    if (e != cudaSuccess) return (int)e;
  }
  int cpg = grid / g->n_graphs;
  if (cpg < 1) cpg = 1;
  int chunk = (g->stride + cpg - 1) / cpg;
  chunk = (chunk + 31) / 32 * 32;
  cpg = (g->stride + chunk - 1) / chunk;
  md_graphs_t gg = *g;
  md_state_t ss = *s;
  int incremental = (ss.lg_scratch != nullptr && !(g_md_debug_flags & (1 << 17))) ? 1 : 0;
  void* args[] = {&gg, &ss, (void*)&forced, &forced_stride, &forced_index, &init, &cpg, &chunk, &incremental};
  e = cudaLaunchCooperativeKernel((const void*)md::env_step_large, dim3(grid), dim3(md::LG_T), args, 0, st);
  if (e != cudaSuccess) return (int)e;
  return 0;
}
