// md_rollout_tail: the END of a greedy rollout, one CTA per graph, everything on chip.
//
// Why: a lock-step batch iteration (md_embed_fused, md_graph_embed, md_qhead, md_env_step) costs ~100 us of fixed latency
// however small the graphs have become, and a dismantling rollout spends MORE THAN HALF of its removals on graphs that
// have already collapsed to a few dozen active nodes (C2: mean 45 removals, the median graph has < 64 active nodes from
// removal ~22 on).  Once a graph has at most TL_N active nodes its whole remaining problem fits one SM's shared memory
// (embeddings of both layers 64 KB, weights 56 KB, alive CSR < 16 KB), so ONE CTA runs forward -> arg-max -> removal ->
// LMCC -> AUDC in a loop until the graph is terminal: no launches, no lock step, no HBM round trips.  The lock-step
// kernels treat a handed-over graph as finished (n_act = 0, done = 1); the tail CTA updates score / traces / counters.
//
// Same arithmetic as the lock-step FP32 kernels it replaces, hence the same reference citations
// (/root/reference/code/MultiDismantler_unit_cost):
//   MultiDismantler_net_graphsage.py:248-321  node features, first embedding, 3 x (spmm n2n, spmm subg, transforms,
//                                             relu, normalise) for node rows and the graph row   (md_embed.cu)
//   MRGNN/mutil_layer_weight.py:266-313, net_graphsage.py:323-393   fusion, Q head, layer gate   (md_qhead.cu)
//   MultiDismantler_torch.py:286-300,769,799-807                    Predict mask + arg-max (ties -> lowest node id)
//   mvc_env.py:74-87,133-162, Mcc.py:3-38, PrepareBatchGraph.py:35-100   removal, LMCC, AUDC, aux features (md_env.cu)
// Two layers only (the reference network); L = 3 batches never hand over.
#include "md_common.cuh"

namespace md {

constexpr int TL_T = 512;
constexpr int TL_W = TL_T / 32;
constexpr int TL_LD = 2 * D + 4;    // final-embedding tile: row = [H_0 (64) | H_1 (64) | pad], as in md_qhead.cu
constexpr int TL_MAP = 2048;        // largest stride for which the global -> local map fits the scratch (X0 doubles as it)

struct TailSmem {
  float A[TL_N * TL_LD];            // embeddings of both layers after the 3 iterations -> T -> e (Q-head tile)
  float X0[TL_N * D];               // ping: current embeddings of the running layer (hand-over: global -> local id map)
  float X1[TL_N * D];               // pong: neighbour sums P, overwritten in place by the new embeddings
  float w13[D * D], w23[D * D], trans[D * D];
  float h1[D * MD_HIDDEN];
  float wn2l[2 * D], bias[D], lw[D], cross[D], h2[MD_HIDDEN + MD_AUX], wl2[MD_GATE];
  float part[TL_T / 64][D];         // column-sum partials
  float ssum[D];                    // per-graph sum of the current embeddings (spmm subgsum)
  float y[2][D], yT[2][D], hid[2][MD_GATE];
  float aux[2][MD_AUX];
  float red_f[TL_W];
  float gc_s[2], gc_g[2], gc_a[2];  // per-graph Q-head constants: s_p, gate softmax, aux . h2[32:]
  float dots[2][2];
  float scr[8];
  unsigned long long red_k[TL_W];
  long long red_ll[TL_W];
  int red_i[TL_W];
  int lab[TL_N], par[TL_N], tmp[TL_N];
  int node[TL_N];                   // row -> graph-local node id (ascending)
  int orig[TL_N];                   // row -> row at hand-over
  int deg[2][TL_N];
  int newidx[TL_N];
  int removed0[TL_N];               // per hand-over row: node removed by the tail
  int node0[TL_N];
  uint16_t rp[2][2][TL_N + 2];      // [buffer][layer] row offsets
  uint8_t col[2][2][TL_E];          // [buffer][layer] neighbour rows
  int s_i[8];
  int st_i[TL_W][2];               // per-warp alive-degree sum / max
  long long st_ll[TL_W];            // per-warp two-hop sum
};
static_assert(sizeof(TailSmem) <= 227 * 1024, "TailSmem exceeds one CTA's shared memory");

__device__ __forceinline__ int tl_find(volatile int* par, int x) {
  int p = par[x];
  while (p != x) {
    const int gp = par[p];
    par[x] = gp;
    x = gp;
    p = par[x];
  }
  return x;
}
__device__ __forceinline__ void tl_unite(int* par, int a, int b) {
  int ra = tl_find(par, a), rb = tl_find(par, b);
  while (ra != rb) {
    const int hi = max(ra, rb), lo = min(ra, rb);
    const int old = atomicCAS(&par[hi], hi, lo);
    if (old == hi) break;
    ra = tl_find(par, old);
    rb = tl_find(par, lo);
  }
}
__device__ __forceinline__ float half_sum16(float v) {  // over the 16 lanes that share a row (tx = lane & 15)
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  return v;
}

constexpr int TL_RG = TL_T / 16;    // rows per row group: thread (tx = tid & 15, ty = tid >> 4) owns rows ty + TL_RG * i, columns 4 tx .. 4 tx + 3
constexpr int TL_NI = TL_N / TL_RG;  // row groups at most

// acc[i][c] += sum_{k < 64} A[(ty + TL_RG i) * lda + k] * W[k * 64 + 4 tx + c]   for the first NI row groups
template <int NI>
__device__ __forceinline__ void tl_gemm64(float (&acc)[TL_NI][4], const float* __restrict__ A, int lda, const float* __restrict__ W,
                                          int tx, int ty) {
#pragma unroll 2
  for (int k0 = 0; k0 < D; k0 += 4) {
    float4 a[NI], w[4];
#pragma unroll
    for (int i = 0; i < NI; ++i) a[i] = *reinterpret_cast<const float4*>(A + (ty + TL_RG * i) * lda + k0);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) w[kk] = *reinterpret_cast<const float4*>(W + (k0 + kk) * D + 4 * tx);
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const float av[4] = {a[i].x, a[i].y, a[i].z, a[i].w};
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        acc[i][0] = fmaf(av[kk], w[kk].x, acc[i][0]);
        acc[i][1] = fmaf(av[kk], w[kk].y, acc[i][1]);
        acc[i][2] = fmaf(av[kk], w[kk].z, acc[i][2]);
        acc[i][3] = fmaf(av[kk], w[kk].w, acc[i][3]);
      }
    }
  }
}

// One message-passing transform for the first NI row groups: out = normalise(relu(P @ w13 + H @ w23)), written (after a
// block barrier: every thread has read its P / H rows) to `out` with row pitch `ldo`.
template <int NI>
__device__ __forceinline__ void tl_transform(TailSmem& sm, const float* P, const float* H, float* out, int ldo, int n, int tx,
                                             int ty) {
  float acc[TL_NI][4];
#pragma unroll
  for (int i = 0; i < NI; ++i)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[i][c] = 0.f;
  tl_gemm64<NI>(acc, P, D, sm.w13, tx, ty);
  tl_gemm64<NI>(acc, H, D, sm.w23, tx, ty);
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      acc[i][c] = fmaxf(acc[i][c], 0.f);
      ss = fmaf(acc[i][c], acc[i][c], ss);
    }
    ss = half_sum16(ss);
    const float inv = __frcp_rn(fmaxf(sqrtf(ss), 1e-12f));  // x / max(|x|, eps) (F.normalize) as x * rcp_rn: <= 1 ulp
    const int r = ty + TL_RG * i;
    if (r < n)
      *reinterpret_cast<float4*>(out + r * ldo + 4 * tx) = make_float4(acc[i][0] * inv, acc[i][1] * inv, acc[i][2] * inv, acc[i][3] * inv);
  }
}

// Fusion of the node rows + Q head for the first NI row groups of the tile sm.A (the body of qhead_kernel<2>, md_qhead.cu).
// Returns this thread's best arg-max key over its rows (valid on tx == 0 lanes).
template <int NI>
__device__ __forceinline__ unsigned long long tl_qhead(TailSmem& sm, int n, int tx, int ty) {
  constexpr int L = 2;
  const float4 bv = *reinterpret_cast<const float4*>(sm.bias + 4 * tx);
  // T_l = tanh(H_l @ trans + bias), in place (a row is private to its half-warp)
#pragma unroll
  for (int l = 0; l < L; ++l) {
    float acc[TL_NI][4];
#pragma unroll
    for (int i = 0; i < NI; ++i)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[i][c] = 0.f;
    tl_gemm64<NI>(acc, sm.A + l * D, TL_LD, sm.trans, tx, ty);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < NI; ++i)
      *reinterpret_cast<float4*>(sm.A + (ty + TL_RG * i) * TL_LD + l * D + 4 * tx) =
          make_float4(tanhf(acc[i][0] + bv.x), tanhf(acc[i][1] + bv.y), tanhf(acc[i][2] + bv.z), tanhf(acc[i][3] + bv.w));
  }
  __syncwarp();
  const float lb = sm.scr[0];
  const float4 lwv = *reinterpret_cast<const float4*>(sm.lw + 4 * tx);
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    float* rowp = sm.A + (ty + TL_RG * i) * TL_LD + 4 * tx;
    float4 Tl[L];
#pragma unroll
    for (int l = 0; l < L; ++l) Tl[l] = *reinterpret_cast<const float4*>(rowp + l * D);
    float dot[L][L];
#pragma unroll
    for (int p = 0; p < L; ++p)
#pragma unroll
      for (int k = 0; k <= p; ++k) {
        float d = Tl[k].x * Tl[p].x * lwv.x;
        d = fmaf(Tl[k].y * Tl[p].y, lwv.y, d);
        d = fmaf(Tl[k].z * Tl[p].z, lwv.z, d);
        d = fmaf(Tl[k].w * Tl[p].w, lwv.w, d);
        d = half_sum16(d);
        dot[p][k] = d;
        dot[k][p] = d;
      }
#pragma unroll
    for (int p = 0; p < L; ++p) {
      float a[L], mx = -1e30f, den = 0.f;
#pragma unroll
      for (int k = 0; k < L; ++k) {
        a[k] = 1.f / (1.f + expf(-(dot[p][k] + lb)));
        mx = fmaxf(mx, a[k]);
      }
#pragma unroll
      for (int k = 0; k < L; ++k) {
        a[k] = expf(a[k] - mx);
        den += a[k];
      }
      float4 x = Tl[p];
#pragma unroll
      for (int k = 0; k < L; ++k)
        if (k != p) {
          const float al = a[k] / den;
          x.x = fmaf(al, Tl[k].x, x.x);
          x.y = fmaf(al, Tl[k].y, x.y);
          x.z = fmaf(al, Tl[k].z, x.z);
          x.w = fmaf(al, Tl[k].w, x.w);
        }
      float ss = x.x * x.x;
      ss = fmaf(x.y, x.y, ss);
      ss = fmaf(x.z, x.z, ss);
      ss = fmaf(x.w, x.w, ss);
      ss = half_sum16(ss);
      const float inv = __frcp_rn(fmaxf(sqrtf(ss), 1e-12f));
      const float s = sm.gc_s[p];
      *reinterpret_cast<float4*>(rowp + p * D) = make_float4(x.x * inv * s, x.y * inv * s, x.z * inv * s, x.w * inv * s);
    }
  }
  __syncwarp();
  float qrow[TL_NI];
#pragma unroll
  for (int i = 0; i < NI; ++i) qrow[i] = 0.f;
  const float2 h2v = *reinterpret_cast<const float2*>(sm.h2 + 2 * tx);
#pragma unroll
  for (int p = 0; p < L; ++p) {
    float acc[TL_NI][2];
#pragma unroll
    for (int i = 0; i < NI; ++i) acc[i][0] = acc[i][1] = 0.f;
#pragma unroll 2
    for (int k0 = 0; k0 < D; k0 += 4) {
      float4 a[NI];
      float2 w[4];
#pragma unroll
      for (int i = 0; i < NI; ++i) a[i] = *reinterpret_cast<const float4*>(sm.A + (ty + TL_RG * i) * TL_LD + p * D + k0);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) w[kk] = *reinterpret_cast<const float2*>(sm.h1 + (k0 + kk) * MD_HIDDEN + 2 * tx);
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const float av[4] = {a[i].x, a[i].y, a[i].z, a[i].w};
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          acc[i][0] = fmaf(av[kk], w[kk].x, acc[i][0]);
          acc[i][1] = fmaf(av[kk], w[kk].y, acc[i][1]);
        }
      }
    }
    const float gw = sm.gc_g[p], aq = sm.gc_a[p];
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      float hq = fmaxf(acc[i][0], 0.f) * h2v.x;
      hq = fmaf(fmaxf(acc[i][1], 0.f), h2v.y, hq);
      hq = half_sum16(hq);
      qrow[i] = fmaf(gw, hq + aq, qrow[i]);
    }
  }
  unsigned long long key = 0ull;
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const int r = ty + TL_RG * i;
    if (r < n && tx == 0) {
      const unsigned long long k2 = ((unsigned long long)orderable(qrow[i]) << 32) | (0xffffffffu - (uint32_t)sm.node[r]);
      key = k2 > key ? k2 : key;
    }
  }
  return key;
}

// block-wide reductions (every thread calls them)
__device__ __forceinline__ int tl_block_sum_i(TailSmem& sm, int v) {
  v = warp_sum_i(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sm.red_i[threadIdx.x >> 5] = v;
  __syncthreads();
  int t = 0;
#pragma unroll
  for (int k = 0; k < TL_W; ++k) t += sm.red_i[k];
  return t;
}
__device__ __forceinline__ int tl_block_max_i(TailSmem& sm, int v) {
  v = warp_max_i(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sm.red_i[threadIdx.x >> 5] = v;
  __syncthreads();
  int t = sm.red_i[0];
#pragma unroll
  for (int k = 1; k < TL_W; ++k) t = max(t, sm.red_i[k]);
  return t;
}
__device__ __forceinline__ long long tl_block_sum_ll(TailSmem& sm, long long v) {
  v = warp_sum_ll(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sm.red_ll[threadIdx.x >> 5] = v;
  __syncthreads();
  long long t = 0;
#pragma unroll
  for (int k = 0; k < TL_W; ++k) t += sm.red_ll[k];
  return t;
}
// sum over the 64 threads tid in [64 l, 64 l + 64) (threads >= 128 pass 0 and ignore the result); block-uniform call
__device__ __forceinline__ float tl_group_sum64(TailSmem& sm, float v) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sm.red_f[threadIdx.x >> 5] = v;
  __syncthreads();
  const int w0 = (threadIdx.x >> 6) * 2;
  return sm.red_f[w0 & (TL_W - 1)] + sm.red_f[(w0 + 1) & (TL_W - 1)];
}

static_assert(TL_N % 32 == 0 && 2 * TL_N <= TL_T && TL_N <= 256, "tail thread layout");

// profiling builds (MDB200_PROFILE=1): per graph {total cycles, steps, nodes at hand-over, SM id} and the phase cycles
// {hand-over setup, embedding iterations, graph rows, Q head + arg-max, LMCC rounds, rest of the env step}
__device__ unsigned long long g_tail_dbg[1024][10];
#ifdef MD_PROFILE
#define TL_CLK(slot)                           \
  do {                                         \
    const long long n_ = clock64();            \
    ph[slot] += n_ - t_ph;                     \
    t_ph = n_;                                 \
  } while (0)
#else
#define TL_CLK(slot) do { } while (0)
#endif

__global__ void __launch_bounds__(TL_T, 1) rollout_tail_kernel(md_graphs_t G, md_state_t S, md_weights_t W) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TailSmem& sm = *reinterpret_cast<TailSmem*>(smem_raw);
  const int g = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tx = tid & 15, ty = tid >> 4;
  constexpr int L = 2;
  const int B = G.n_graphs, St = G.stride, base = g * St;

  // ---- claim the graph (tail_n[g] = number of active nodes at hand-over; 0 = nothing to do) ---------------------
  if (tid == 0) sm.s_i[0] = atomicExch(&S.tail_n[g], 0);
  __syncthreads();
  int n = sm.s_i[0];
  if (n <= 0 || n > TL_N) return;
  const int n0 = n;
#ifdef MD_PROFILE
  long long ph[6] = {0, 0, 0, 0, 0, 0};
  const long long t_begin = clock64();
  long long t_ph = t_begin;
#endif
  const int N = G.n_nodes[g];
  const bool degree_cost = S.cost_mode != 0;

  // ---- weights into shared memory ----------------------------------------------------------------------------------
  for (int i = tid; i < D * D / 4; i += TL_T) {
    reinterpret_cast<float4*>(sm.w13)[i] = reinterpret_cast<const float4*>(W.w13)[i];
    reinterpret_cast<float4*>(sm.w23)[i] = reinterpret_cast<const float4*>(W.w23)[i];
    reinterpret_cast<float4*>(sm.trans)[i] = reinterpret_cast<const float4*>(W.trans)[i];
  }
  for (int i = tid; i < D * MD_HIDDEN / 4; i += TL_T) reinterpret_cast<float4*>(sm.h1)[i] = reinterpret_cast<const float4*>(W.h1)[i];
  if (tid < 2 * D) sm.wn2l[tid] = W.w_n2l[tid];
  if (tid < D) {
    sm.bias[tid] = W.bias[tid];
    sm.lw[tid] = W.logis_w[tid];
    sm.cross[tid] = W.cross[tid];
  }
  if (tid < MD_HIDDEN + MD_AUX) sm.h2[tid] = W.h2[tid];
  if (tid < MD_GATE) sm.wl2[tid] = W.wl2[tid];
  if (tid == 0) sm.scr[0] = W.logis_b[0];

  // ---- graph state at hand-over -----------------------------------------------------------------------------------
  int* map = reinterpret_cast<int*>(sm.X0);  // graph-local node id -> row (only rows of active nodes are read back)
  for (int r = tid; r < n; r += TL_T) {
    const int v = S.act_list[base + r] - base;
    sm.node[r] = v;
    sm.node0[r] = v;
    sm.orig[r] = r;
    sm.removed0[r] = 0;
    map[v] = r;
  }
  __syncthreads();
  for (int r = tid; r < n; r += TL_T) {
    const int v = sm.node[r];
    sm.lab[r] = map[S.label[base + v] - base];  // label = smallest member of the mutual component: an active node
    sm.deg[0][r] = S.deg[0][base + v];
    sm.deg[1][r] = S.deg[1][base + v];
  }
  for (int l = 0; l < L; ++l) {
    const int32_t* __restrict__ ap = S.aptr[l] + base;
    const int32_t* __restrict__ ac = S.acol[l];
    const int e0 = ap[0], cnt = ap[n] - e0;
    for (int r = tid; r <= n; r += TL_T) sm.rp[0][l][r] = (uint16_t)(ap[r] - e0);
    for (int e = tid; e < cnt; e += TL_T) sm.col[0][l][e] = (uint8_t)map[ac[e0 + e] - base];
  }
  if (tid < L * MD_AUX) sm.aux[tid / MD_AUX][tid % MD_AUX] = S.aux[((size_t)g * L + tid / MD_AUX) * MD_AUX + tid % MD_AUX];
  __syncthreads();  // (map is dead from here on: X0 becomes an embedding buffer)
  int cur = 0;      // CSR buffer in use
  int dmax[L] = {S.max_deg[0 * B + g], S.max_deg[1 * B + g]};
  int n_removed = S.n_removed[g];
  int cov[L] = {S.cov_edges[0 * B + g], S.cov_edges[1 * B + g]};
  int n_steps = S.n_steps[g];
  double score = S.score[g];
  int rank = S.rank[g];
  const int E[L] = {G.n_edges[0 * B + g], G.n_edges[1 * B + g]};
  const double mr = (double)G.max_rank[g];
  bool terminal = false;

  TL_CLK(0);
  for (int guard = 0; guard <= n0 && !terminal; ++guard) {
    const int ni = (n + TL_RG - 1) / TL_RG;  // row groups in use
    // ================================ forward ================================
    for (int l = 0; l < L; ++l) {
      // node features + first embedding (md_embed.cu embed_init_kernel): one warp per row, 2 columns per lane
      {
        const float w0a = sm.wn2l[2 * lane], w0b = sm.wn2l[2 * lane + 1];
        const float w1a = sm.wn2l[D + 2 * lane], w1b = sm.wn2l[D + 2 * lane + 1];
        const float dm = (float)dmax[l];
        for (int r = warp; r < n; r += TL_W) {
          float x0, x1;
          if (degree_cost) {
            x0 = G.node_w[l][base + sm.node[r]];
            x1 = 1.0f;
          } else {
            x0 = (float)sm.deg[l][r] / dm;
            x1 = x0;
          }
          const float a = fmaxf(fmaf(x1, w1a, x0 * w0a), 0.f);
          const float b = fmaxf(fmaf(x1, w1b, x0 * w0b), 0.f);
          const float den = fmaxf(sqrtf(warp_sum(a * a + b * b)), 1e-12f);
          *reinterpret_cast<float2*>(sm.X0 + r * D + 2 * lane) = make_float2(a / den, b / den);
        }
        // graph row: y0 = normalise(relu([1, 1] @ w_n2l))
        float yv = 0.f;
        if (tid < D) yv = fmaxf(sm.wn2l[tid] + sm.wn2l[D + tid], 0.f);
        const float nrm = sqrtf(tl_group_sum64(sm, yv * yv));  // (block barrier inside: X0 rows are complete after it)
        if (tid < D) sm.y[l][tid] = yv / fmaxf(nrm, 1e-12f);
        __syncthreads();
      }
      float* Hc = sm.X0;  // current embeddings
      float* Hn = sm.X1;  // neighbour sums, then the new embeddings
      for (int it = 0; it < MD_BP_ITERS; ++it) {
        // (a) per-graph column sums of the current embeddings (spmm subgsum): TL_T / 64 row groups x 64 columns, fixed order
        {
          const int c = tid & 63, q = tid >> 6;
          float s = 0.f;
          for (int r = q; r < n; r += TL_T / 64) s += Hc[r * D + c];
          sm.part[q][c] = s;
        }
        __syncthreads();
        if (tid < D) {
          float s = 0.f;
#pragma unroll
          for (int q = 0; q < TL_T / 64; ++q) s += sm.part[q][tid];
          sm.ssum[tid] = s;
        }
        __syncthreads();
        // (b) warps 0, 1: graph row  y' = normalise(relu(S @ w13 + y @ w23));  the other warps: P = A_alive . H, CSR order
        float yo = 0.f;
        if (warp < 2) {
          const int c = tid;
          float o = 0.f;
#pragma unroll 8
          for (int k = 0; k < D; ++k) o = fmaf(sm.ssum[k], sm.w13[k * D + c], o);
#pragma unroll 8
          for (int k = 0; k < D; ++k) o = fmaf(sm.y[l][k], sm.w23[k * D + c], o);
          yo = fmaxf(o, 0.f);
        } else {
          const uint16_t* rp = sm.rp[cur][l];
          const uint8_t* cl = sm.col[cur][l];
          for (int r = warp - 2; r < n; r += TL_W - 2) {
            float2 acc = make_float2(0.f, 0.f);
            const int e1 = rp[r + 1];
            for (int e = rp[r]; e < e1; ++e) {
              const float2 h = *reinterpret_cast<const float2*>(Hc + (int)cl[e] * D + 2 * lane);
              acc.x += h.x;
              acc.y += h.y;
            }
            *reinterpret_cast<float2*>(Hn + r * D + 2 * lane) = acc;
          }
        }
        {
          const float nrm = sqrtf(tl_group_sum64(sm, warp < 2 ? yo * yo : 0.f));  // (block barriers: P complete, y reads done)
          if (warp < 2) sm.y[l][tid] = yo / fmaxf(nrm, 1e-12f);
        }
        // (c) node rows: H' = normalise(relu(P @ w13 + H @ w23)); the last iteration writes the Q-head tile directly
        const bool last = it + 1 == MD_BP_ITERS;
        float* out = last ? sm.A + l * D : Hn;
        const int ldo = last ? TL_LD : D;
        static_assert(TL_NI == 2, "dispatch below assumes two row groups");
        if (ni <= 1) tl_transform<1>(sm, Hn, Hc, out, ldo, n, tx, ty);
        else tl_transform<2>(sm, Hn, Hc, out, ldo, n, tx, ty);
        __syncthreads();
        float* t = Hc;
        Hc = Hn;
        Hn = t;
      }
    }
    TL_CLK(1);
    // ---- graph rows: fusion (BitwiseMultipyLogis), s_p = yf_p . cross, layer gate, aux . h2 (md_embed.cu graph_embed_kernel) ----
    {
      const int l = (tid >> 6) & 1, c = tid & 63;
      const bool act = tid < L * D;
      float o = 0.f;
      if (act) {
#pragma unroll 8
        for (int k = 0; k < D; ++k) o = fmaf(sm.y[l][k], sm.trans[k * D + c], o);
        sm.yT[l][c] = tanhf(o + sm.bias[c]);
      }
      __syncthreads();
      const float lwc = sm.lw[c], lb = sm.scr[0];
      for (int k = 0; k < L; ++k) {
        const float d = tl_group_sum64(sm, act ? sm.yT[k][c] * sm.yT[l][c] * lwc : 0.f);
        if (act && c == 0) sm.dots[l][k] = d;
      }
      __syncthreads();
      float x = act ? sm.yT[l][c] : 0.f;
      if (act) {
        float a[L], mx = -1e30f, den = 0.f;
        for (int k = 0; k < L; ++k) {
          a[k] = 1.f / (1.f + expf(-(sm.dots[l][k] + lb)));
          mx = fmaxf(mx, a[k]);
        }
        for (int k = 0; k < L; ++k) {
          a[k] = expf(a[k] - mx);
          den += a[k];
        }
        for (int k = 0; k < L; ++k)
          if (k != l) x = fmaf(a[k] / den, sm.yT[k][c], x);
      }
      {
        const float nrm = sqrtf(tl_group_sum64(sm, x * x));
        x = x / fmaxf(nrm, 1e-12f);
      }
      __syncthreads();
      if (act) sm.y[l][c] = x;  // yf_l
      {
        const float s = tl_group_sum64(sm, act ? x * sm.cross[c] : 0.f);  // (barriers inside: yf is visible afterwards)
        if (act && c == 0) sm.gc_s[l] = s;
      }
      {  // gate_l = relu(yf_l @ w_layer1) @ w_layer2 : thread (l2, j) for j < 128
        const int l2 = (tid >> 7) & 1, j = tid & 127;
        if (tid < L * MD_GATE) {
          float o2 = 0.f;
#pragma unroll 16
          for (int k = 0; k < D; ++k) o2 = fmaf(sm.y[l2][k], __ldg(W.wl1 + k * MD_GATE + j), o2);
          sm.hid[l2][j] = fmaxf(o2, 0.f) * sm.wl2[j];
        }
      }
      __syncthreads();
      {
        const float gsum = tl_group_sum64(sm, act ? sm.hid[l][c] + sm.hid[l][c + D] : 0.f);
        if (act && c == 0) sm.scr[2 + l] = gsum;
      }
      __syncthreads();
      if (tid == 0) {
        const float g0 = sm.scr[2], g1 = sm.scr[3];
        const float mx = fmaxf(g0, g1);
        const float e0 = expf(g0 - mx), e1 = expf(g1 - mx);
        const float den = e0 + e1;
        sm.gc_g[0] = e0 / den;
        sm.gc_g[1] = e1 / den;
        for (int k = 0; k < L; ++k) {
          float aq = 0.f;
          for (int i = 0; i < MD_AUX; ++i) aq = fmaf(sm.aux[k][i], sm.h2[MD_HIDDEN + i], aq);
          sm.gc_a[k] = aq;
        }
      }
      __syncthreads();
    }
    TL_CLK(2);
    // ---- node rows: fusion + Q head + arg-max ----------------------------------------------------------------
    unsigned long long key;
    if (ni <= 1) key = tl_qhead<1>(sm, n, tx, ty);
    else key = tl_qhead<2>(sm, n, tx, ty);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
      key = other > key ? other : key;
    }
    __syncthreads();
    if (lane == 0) sm.red_k[warp] = key;
    __syncthreads();
    if (tid == 0) {
      unsigned long long k = sm.red_k[0];
      for (int w2 = 1; w2 < TL_W; ++w2) k = sm.red_k[w2] > k ? sm.red_k[w2] : k;
      const int a_node = (int)(0xffffffffu - (uint32_t)(k & 0xffffffffull));
      int a_row = -1;
      for (int r = 0; r < n; ++r)
        if (sm.node[r] == a_node) a_row = r;
      sm.s_i[1] = a_row;
      sm.s_i[2] = a_node;
    }
    __syncthreads();
    const int a_row = sm.s_i[1], a_node = sm.s_i[2];
    if (a_row < 0) break;  // (cannot happen: n >= 2 active rows always yield a key)

    TL_CLK(3);
    // ================================ environment step ================================
    // numCoveredEdges[l] += alive degree of a (mvc_env.py:81-84); remove a
    cov[0] += sm.deg[0][a_row];
    cov[1] += sm.deg[1][a_row];
    n_removed += 1;
    __syncthreads();
    if (tid == 0) {
      sm.lab[a_row] = -1;
      sm.removed0[sm.orig[a_row]] = 1;
    }
    __syncthreads();
    // Mcc.MCC: refine the partition layer by layer until both layers are stable (labels = smallest row of the block)
    {
      int stable = 0, l = 0;
      while (stable < L) {
        for (int r = tid; r < n; r += TL_T) sm.par[r] = r;
        __syncthreads();
        const uint16_t* rp = sm.rp[cur][l];
        const uint8_t* cl = sm.col[cur][l];
        for (int r = tid; r < n; r += TL_T) {
          const int lr = sm.lab[r];
          if (lr < 0) continue;
          const int e1 = rp[r + 1];
          for (int e = rp[r]; e < e1; ++e) {
            const int u = cl[e];
            if (u < r && sm.lab[u] == lr) tl_unite(sm.par, u, r);
          }
        }
        __syncthreads();
        int changed = 0;
        for (int r = tid; r < n; r += TL_T) {
          int root = -1;
          if (sm.lab[r] >= 0) {
            root = tl_find(sm.par, r);
            changed |= (root != sm.lab[r]);
          }
          sm.tmp[r] = root;
        }
        changed = __syncthreads_or(changed);
        for (int r = tid; r < n; r += TL_T) sm.lab[r] = sm.tmp[r];
        __syncthreads();
        stable = changed ? 1 : stable + 1;
        l ^= 1;
      }
    }
    TL_CLK(4);
    // component sizes -> LMCC rank (every node outside the tail's rows is removed or a singleton)
    for (int r = tid; r < n; r += TL_T) sm.tmp[r] = 0;
    __syncthreads();
    for (int r = tid; r < n; r += TL_T)
      if (sm.lab[r] >= 0) atomicAdd(&sm.tmp[sm.lab[r]], 1);
    __syncthreads();
    {
      int mx = 0;
      for (int r = tid; r < n; r += TL_T) mx = max(mx, sm.tmp[r]);
      rank = tl_block_max_i(sm, mx);
      if (rank < 1 && N - n_removed > 0) rank = 1;
    }
    // alive degrees, status counters (get_status_info): thread (l, r)
    int alive[L];
    long long twohop[L];
    {
      const int l = (tid / TL_N) & 1, r = tid % TL_N;  // threads [0, TL_N): layer 0, [TL_N, 2 TL_N): layer 1 (whole warps per layer)
      int d = 0;
      if (tid < L * TL_N && r < n && sm.lab[r] >= 0) {
        const uint16_t* rp = sm.rp[cur][l];
        const uint8_t* cl = sm.col[cur][l];
        const int lr = sm.lab[r], e1 = rp[r + 1];
        for (int e = rp[r]; e < e1; ++e) d += (sm.lab[cl[e]] == lr);
      }
      if (tid < L * TL_N && r < n) sm.deg[l][r] = d;
      const int ws = warp_sum_i(d), wm = warp_max_i(d);
      const long long wt = warp_sum_ll((long long)d * (d - 1) / 2);
      __syncthreads();
      if (lane == 0) {
        sm.st_i[warp][0] = ws;
        sm.st_i[warp][1] = wm;
        sm.st_ll[warp] = wt;
      }
      __syncthreads();
      constexpr int WPL = TL_N / 32;  // warps per layer
#pragma unroll
      for (int ll = 0; ll < L; ++ll) {
        int a = 0, m = 0;
        long long t = 0;
#pragma unroll
        for (int k = 0; k < WPL; ++k) {
          a += sm.st_i[ll * WPL + k][0];
          m = max(m, sm.st_i[ll * WPL + k][1]);
          t += sm.st_ll[ll * WPL + k];
        }
        alive[ll] = a;
        dmax[ll] = m;
        twohop[ll] = t;
      }
    }
    terminal = (alive[0] == 0) || (alive[1] == 0);
    // reward / AUDC in f64 with the reference's operation order (md_env.cu)
    {
      const double rk = (double)rank;
      double r, mcc;
      if (!degree_cost) {
        r = __ddiv_rn(rk, __dmul_rn(mr, (double)N));
        mcc = __dmul_rn(r, (double)N);
      } else {
        mcc = __ddiv_rn(rk, mr);
        r = __ddiv_rn(__dmul_rn(mcc, G.node_cost[base + a_node]), 2.0);
      }
      score = __dadd_rn(score, r);
      if (tid == 0 && n_steps < S.max_steps) {
        S.trace_action[(size_t)g * S.max_steps + n_steps] = a_node;
        S.trace_rank[(size_t)g * S.max_steps + n_steps] = rank;
        S.trace_maxcc[(size_t)g * S.max_steps + n_steps] = mcc;
      }
      n_steps += 1;
    }
    if (tid < L) {  // aux features of the next forward (PrepareBatchGraph.py:92-100): f64 quotients cast to f32
      const int ll = tid;
      sm.aux[ll][0] = (float)__ddiv_rn((double)n_removed, (double)N);
      sm.aux[ll][1] = E[ll] > 0 ? (float)__ddiv_rn((double)cov[ll], (double)E[ll]) : 0.f;
      sm.aux[ll][2] = (float)__ddiv_rn((double)twohop[ll], (double)((long long)N * (long long)N));
      sm.aux[ll][3] = 1.0f;
    }
    if (terminal) break;
    // ---- compaction: rows that are still active (alive degree > 0), ascending; alive CSR with renumbered rows ----
    {
      const int r = tid;
      const int flag = (r < n && sm.deg[0][r] > 0) ? 1 : 0;
      const unsigned bal = __ballot_sync(0xffffffffu, flag);
      __syncthreads();
      if (lane == 0) sm.red_i[warp] = __popc(bal);
      __syncthreads();
      int woff = 0, total = 0;
#pragma unroll
      for (int k = 0; k < TL_W; ++k) {
        const int c = sm.red_i[k];
        if (k < warp) woff += c;
        total += c;
      }
      const int q = woff + __popc(bal & ((1u << lane) - 1u));
      if (r < n) sm.newidx[r] = flag ? q : -1;
      __syncthreads();
      const int n_new = total;
      // row offsets of the new CSR: exclusive scan of the alive degrees of the kept rows, per layer (warp l scans)
      if (warp < L) {
        const int l = warp;
        int running = 0;
        for (int r0 = 0; r0 < n; r0 += 32) {
          const int rr = r0 + lane;
          const int keep = rr < n && sm.newidx[rr] >= 0;
          const int d = keep ? sm.deg[l][rr] : 0;
          int incl = d;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int y2 = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y2;
          }
          if (keep) sm.rp[cur ^ 1][l][sm.newidx[rr]] = (uint16_t)(running + incl - d);
          running += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) sm.rp[cur ^ 1][l][n_new] = (uint16_t)running;
      }
      __syncthreads();
      {
        const int l = (tid / TL_N) & 1, rr = tid % TL_N;
        if (tid < L * TL_N && rr < n && sm.newidx[rr] >= 0) {
          const uint16_t* rp = sm.rp[cur][l];
          const uint8_t* cl = sm.col[cur][l];
          uint8_t* out = sm.col[cur ^ 1][l];
          int o = sm.rp[cur ^ 1][l][sm.newidx[rr]];
          const int lr = sm.lab[rr], e1 = rp[rr + 1];
          for (int e = rp[rr]; e < e1; ++e) {
            const int u = cl[e];
            if (sm.lab[u] == lr) out[o++] = (uint8_t)sm.newidx[u];
          }
        }
      }
      // per-row arrays: gather through registers, then scatter to the new positions
      int v_node = 0, v_orig = 0, v_lab = 0, v_d0 = 0, v_d1 = 0, v_new = -1;
      if (tid < n) {
        v_new = sm.newidx[tid];
        if (v_new >= 0) {
          v_node = sm.node[tid];
          v_orig = sm.orig[tid];
          v_lab = sm.newidx[sm.lab[tid]];  // the smallest row of a block of >= 2 nodes is active
          v_d0 = sm.deg[0][tid];
          v_d1 = sm.deg[1][tid];
        }
      }
      __syncthreads();
      if (v_new >= 0) {
        sm.node[v_new] = v_node;
        sm.orig[v_new] = v_orig;
        sm.lab[v_new] = v_lab;
        sm.deg[0][v_new] = v_d0;
        sm.deg[1][v_new] = v_d1;
      }
      __syncthreads();
      cur ^= 1;
      n = n_new;
    }
    TL_CLK(5);
  }
#ifdef MD_PROFILE
  if (tid == 0 && g < 1024) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    g_tail_dbg[g][0] = (unsigned long long)(clock64() - t_begin);
    g_tail_dbg[g][1] = (unsigned long long)(n_steps - S.n_steps[g]);
    g_tail_dbg[g][2] = (unsigned long long)n0;
    g_tail_dbg[g][3] = smid;
    for (int i = 0; i < 6; ++i) g_tail_dbg[g][4 + i] = (unsigned long long)ph[i];
  }
#endif

  // ---- write the final state back (the graph is terminal: every remaining node is a singleton) ----------------------
  for (int r = tid; r < n0; r += TL_T) {
    const int v = base + sm.node0[r];
    S.label[v] = sm.removed0[r] ? -1 : v;
    S.deg[0][v] = 0;
    S.deg[1][v] = 0;
  }
  if (tid == 0) {
    S.score[g] = score;
    S.n_steps[g] = n_steps;
    S.rank[g] = rank;
    S.n_removed[g] = n_removed;
    for (int l = 0; l < L; ++l) {
      S.cov_edges[l * B + g] = cov[l];
      S.alive_edges[l * B + g] = 0;
      S.max_deg[l * B + g] = 0;
      S.twohop[l * B + g] = 0;
      for (int i = 0; i < MD_AUX; ++i) S.aux[((size_t)g * L + l) * MD_AUX + i] = sm.aux[l][i];
    }
    S.n_act[g] = 0;
    S.done[g] = 1;
    __threadfence();
    atomicAdd(S.n_done, 1);
  }
}

// Hand-over test at the start of a rollout (after md_env_reset): graphs that are small enough from the start.
__global__ void tail_handover_kernel(md_graphs_t G, md_state_t S) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G.n_graphs || S.done[g]) return;
  const int n = S.n_act[g];
  if (n <= 0 || n > TL_N) return;
  int worst = 0;
  for (int l = 0; l < G.n_layers; ++l) worst = max(worst, 2 * S.alive_edges[l * G.n_graphs + g]);
  if (worst > TL_E) return;
  S.tail_n[g] = n;
  S.n_act[g] = 0;
  S.done[g] = 1;
  atomicAdd(&S.flags[FLAG_LEFT], 1);
}

}  // namespace md

extern "C" int md_debug_read_tail(unsigned long long* host10240) {
  return (int)cudaMemcpyFromSymbol(host10240, md::g_tail_dbg, sizeof(unsigned long long) * 10240);
}
extern "C" int32_t md_tail_max_nodes(void) { return md::TL_N; }
extern "C" int32_t md_tail_max_entries(void) { return md::TL_E; }

// 1 if graphs of this batch may be handed over to md_rollout_tail
extern "C" int32_t md_tail_applies(const md_graphs_t* g) {
  return g && g->n_layers == 2 && g->stride <= md::TL_MAP && md_env_uses_small_path(g);
}

extern "C" int md_tail_handover(const md_graphs_t* g, const md_state_t* s, md_stream_t stream) {
  if (!g || !s || !s->tail_n || !s->flags) return MD_EINVAL;
  if (!md_tail_applies(g)) return MD_ETOOBIG;
  md::tail_handover_kernel<<<(g->n_graphs + 127) / 128, 128, 0, (cudaStream_t)stream>>>(*g, *s);
  MD_CHECK_LAUNCH();
  return 0;
}

// ---- the tail stream: low priority, one per device ------------------------------------------------------------------
namespace {
constexpr int TAIL_STREAMS = 4;  // round robin: a launch whose CTAs run long must not hold back the next hand-overs
struct TailCtx {
  cudaStream_t st[TAIL_STREAMS] = {};
  cudaEvent_t ev_fork = nullptr, ev_join[TAIL_STREAMS] = {};
  unsigned next = 0;
};
md::PerDevice<TailCtx> g_tail_ctx;
int tail_ctx(TailCtx** out) {
  TailCtx& c = g_tail_ctx.get();
  if (!c.ev_fork) {
    int least = 0, greatest = 0;
    cudaError_t e = cudaDeviceGetStreamPriorityRange(&least, &greatest);
    if (e != cudaSuccess) return (int)e;
    for (int i = 0; i < TAIL_STREAMS; ++i) {
      e = cudaStreamCreateWithPriority(&c.st[i], cudaStreamNonBlocking, least);
      if (e != cudaSuccess) return (int)e;
      e = cudaEventCreateWithFlags(&c.ev_join[i], cudaEventDisableTiming);
      if (e != cudaSuccess) return (int)e;
    }
    e = cudaEventCreateWithFlags(&c.ev_fork, cudaEventDisableTiming);
    if (e != cudaSuccess) return (int)e;
  }
  *out = &c;
  return 0;
}
}  // namespace

extern "C" int md_rollout_tail(const md_graphs_t* g, const md_state_t* s, const md_weights_t* w, md_stream_t stream) {
  if (!g || !s || !w || !s->tail_n) return MD_EINVAL;
  if (!md_tail_applies(g)) return MD_ETOOBIG;
  static md::PerDevice<bool> configured_pd;
  bool& configured = configured_pd.get();
  const int smem = (int)sizeof(md::TailSmem);
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(md::rollout_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  TailCtx* c = nullptr;
  int rc = tail_ctx(&c);
  if (rc) return rc;
  // fork: the tail launch sees everything enqueued on `stream` so far, `stream` itself does not wait for it
  cudaStream_t ts = c->st[c->next++ % TAIL_STREAMS];
  cudaError_t e = cudaEventRecord(c->ev_fork, (cudaStream_t)stream);
  if (e != cudaSuccess) return (int)e;
  e = cudaStreamWaitEvent(ts, c->ev_fork, 0);
  if (e != cudaSuccess) return (int)e;
  md::rollout_tail_kernel<<<g->n_graphs, md::TL_T, smem, ts>>>(*g, *s, *w);
  MD_CHECK_LAUNCH();
  return 0;
}

extern "C" int md_tail_join(md_stream_t stream) {
  TailCtx* c = nullptr;
  int rc = tail_ctx(&c);
  if (rc) return rc;
  for (int i = 0; i < TAIL_STREAMS; ++i) {
    cudaError_t e = cudaEventRecord(c->ev_join[i], c->st[i]);
    if (e != cudaSuccess) return (int)e;
    e = cudaStreamWaitEvent((cudaStream_t)stream, c->ev_join[i], 0);
    if (e != cudaSuccess) return (int)e;
  }
  return 0;
}
