// Node-embedding kernels of the Q network.
//
// Replaces (reference, /root/reference/code/MultiDismantler_unit_cost/MultiDismantler_net_graphsage.py):
//   md_embed_init  : :248-287   node features (alive degree / max alive degree, or the degree-cost weights of
//                               D/PrepareBatchGraph.py:133-136), relu(x @ w_n2l), L2 normalise
//   md_embed_iter  : :288-310   torch_sparse.spmm(n2nsum) + three dense transforms + relu + normalise, and the
//                               torch_sparse.spmm(subgsum) of :297-298 as per-tile partial sums
//   md_graph_embed : :272-287, :301-321 (virtual node chain), U/MRGNN/mutil_layer_weight.py:266-313 (fusion of the
//                               graph rows), net_graphsage.py:350-351,377-390 (per-graph parts of the Q head)
//
// Layout: embeddings H[l] are [n_total][64] f32 rows addressed by global node id, so neighbour gathers need no
// re-indexing; only the rows listed in act_list (active nodes, compacted per graph by the env kernel) are touched.
#include "md_common.cuh"

namespace md {

// =================================================================================================
// md_embed_init
// =================================================================================================
__global__ void __launch_bounds__(256) embed_init_kernel(md_graphs_t G, md_state_t S, md_weights_t W, md_work_t K) {
  const int l = blockIdx.y;
  const int tpg = G.stride / TILE_M;
  const int g = blockIdx.x / tpg;
  const int t = blockIdx.x % tpg;
  const int nact = S.n_act[g];
  if (t * TILE_M >= nact) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float w0a = W.w_n2l[2 * lane], w0b = W.w_n2l[2 * lane + 1];
  const float w1a = W.w_n2l[D + 2 * lane], w1b = W.w_n2l[D + 2 * lane + 1];
  const bool degree_cost = G.node_w[0] != nullptr;
  const float dmax = (float)S.max_deg[l * G.n_graphs + g];
  float* __restrict__ H = K.h[0][l];
  for (int r = t * TILE_M + warp; r < min(nact, (t + 1) * TILE_M); r += 8) {
    const int v = S.act_list[g * G.stride + r];
    float x0, x1;
    if (degree_cost) {
      x0 = G.node_w[l][v];
      x1 = 1.0f;
    } else {
      x0 = (float)S.deg[l][v] / dmax;  // degree / degree_max, both exact small integers in f32 (:255-258)
      x1 = x0;
    }
    float a = fmaxf(fmaf(x1, w1a, x0 * w0a), 0.f);
    float b = fmaxf(fmaf(x1, w1b, x0 * w0b), 0.f);
    const float nrm = sqrtf(warp_sum(a * a + b * b));
    const float den = fmaxf(nrm, 1e-12f);  // F.normalize eps
    *reinterpret_cast<float2*>(H + (size_t)v * D + 2 * lane) = make_float2(a / den, b / den);
  }
}

// =================================================================================================
// psum_reduce : graphs of many tiles (N = 1M: 7813 per layer) -- the per-tile column sums of one (iteration, layer, graph)
// are folded into the slot of tile 0 by a whole CTA (16 tile groups x 64 columns, 8 independent partial sums per thread,
// combined in a fixed order), so that graph_embed_kernel's single thread per column does not walk 3 x 7813 dependent
// loads (0.88 ms per step at C4).  Graphs of few tiles keep the plain sequential sum inside graph_embed_kernel.
// =================================================================================================
constexpr int PSUM_REDUCE_MIN_TILES = 64;

__global__ void __launch_bounds__(1024) psum_reduce_kernel(md_graphs_t G, md_state_t S, md_work_t K) {
  __shared__ float part[16][D];
  const int L = G.n_layers, B = G.n_graphs;
  const int g = blockIdx.x % B, l = (blockIdx.x / B) % L, it = blockIdx.x / (B * L);
  const int tpg = G.stride / TILE_M, tiles_per_layer = B * tpg;
  const int ntile = (S.n_act[g] + TILE_M - 1) / TILE_M;
  if (ntile <= PSUM_REDUCE_MIN_TILES) return;
  const int c = threadIdx.x & 63, q = threadIdx.x >> 6;
  float* ps = K.psum + (((size_t)it * L + l) * tiles_per_layer + (size_t)g * tpg) * D + c;
  float a[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = 0.f;
  for (int t0 = q; t0 < ntile; t0 += 16 * 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int t = t0 + 16 * j;
      if (t < ntile) a[j] += ps[(size_t)t * D];
    }
  }
  part[q][c] = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
  __syncthreads();
  if (q == 0) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += part[k][c];
    ps[0] = s;
  }
}

// =================================================================================================
// md_graph_embed : one CTA per graph, L*64 threads
// =================================================================================================
__device__ __forceinline__ float group_sum64(float v, float* scratch, int l, int c) {
  // sum over the 64 threads (2 warps) of layer l
  v = warp_sum(v);
  __syncthreads();
  if ((c & 31) == 0) scratch[l * 2 + (c >> 5)] = v;
  __syncthreads();
  return scratch[l * 2] + scratch[l * 2 + 1];
}

__global__ void __launch_bounds__(MD_MAX_LAYERS* D) graph_embed_kernel(md_graphs_t G, md_state_t S, md_weights_t W,
                                                                      md_work_t K) {
  const int g = blockIdx.x;
  const int L = G.n_layers;
  const int l = threadIdx.x >> 6, c = threadIdx.x & 63;
  __shared__ float y[MD_MAX_LAYERS][D], sbuf[MD_MAX_LAYERS][D], T[MD_MAX_LAYERS][D], hid[MD_MAX_LAYERS][MD_GATE];
  __shared__ float scratch[2 * MD_MAX_LAYERS];
  __shared__ float dots[MD_MAX_LAYERS][MD_MAX_LAYERS], sval[MD_MAX_LAYERS], gate[MD_MAX_LAYERS];
  const int nact = S.n_act[g];
  const int tpg = G.stride / TILE_M;
  const int ntile = (nact + TILE_M - 1) / TILE_M;
  const int tiles_per_layer = G.n_graphs * tpg;

  // y0 = normalize(relu([1,1] @ w_n2l))   (:272-287)
  float yv = fmaxf(W.w_n2l[c] + W.w_n2l[D + c], 0.f);
  {
    const float nrm = sqrtf(group_sum64(yv * yv, scratch, l, c));
    yv = yv / fmaxf(nrm, 1e-12f);
  }
  y[l][c] = yv;
  // The kernel is a chain of small dependent steps: everything that does not depend on the chain is fetched up front, with
  // all loads in flight at once -- this thread's columns of W13 / W23 (reused by the three iterations) and of trans, and the
  // per-tile column sums of all three iterations (before: ~9 exposed L2 round-trip chains, 16 us per launch).
  float w13c[D], w23c[D], trc[D], sv[MD_BP_ITERS];
#pragma unroll
  for (int k = 0; k < D; ++k) w13c[k] = W.w13[k * D + c], w23c[k] = W.w23[k * D + c];
#pragma unroll
  for (int k = 0; k < D; ++k) trc[k] = W.trans[k * D + c];
  {
    const int nsum = ntile > PSUM_REDUCE_MIN_TILES ? 1 : ntile;  // many tiles: psum_reduce_kernel left the total in tile 0
    // all loads of a group of 8 tiles (and of all three iterations) in flight before the first add: as a plain loop the
    // sums were a chain of up to 8 L2 round trips per iteration -- 29 % of the kernel's stall samples.  Same order of adds.
#pragma unroll
    for (int it = 0; it < MD_BP_ITERS; ++it) sv[it] = 0.f;
    for (int t0 = 0; t0 < nsum; t0 += 8) {
      float pv[MD_BP_ITERS][8];
#pragma unroll
      for (int it = 0; it < MD_BP_ITERS; ++it) {
        const float* ps = K.psum + (((size_t)it * L + l) * tiles_per_layer + (size_t)g * tpg) * D + c;
#pragma unroll
        for (int j = 0; j < 8; ++j) pv[it][j] = t0 + j < nsum ? ps[(size_t)(t0 + j) * D] : 0.f;
      }
#pragma unroll
      for (int it = 0; it < MD_BP_ITERS; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (t0 + j < nsum) sv[it] += pv[it][j];
      }
    }
  }
#pragma unroll
  for (int it = 0; it < MD_BP_ITERS; ++it) {
    sbuf[l][c] = sv[it];
    __syncthreads();
    float o = 0.f;
#pragma unroll
    for (int k = 0; k < D; ++k) o = fmaf(sbuf[l][k], w13c[k], o);
#pragma unroll
    for (int k = 0; k < D; ++k) o = fmaf(y[l][k], w23c[k], o);
    o = fmaxf(o, 0.f);
    const float nrm = sqrtf(group_sum64(o * o, scratch, l, c));  // contains __syncthreads: y reads are done
    y[l][c] = o / fmaxf(nrm, 1e-12f);
    __syncthreads();
  }

  // (the two columns of w_layer1 this thread needs at the very end: in flight under the fusion phase)
  float wg0[D], wg1[D];
#pragma unroll
  for (int k = 0; k < D; ++k) wg0[k] = W.wl1[k * MD_GATE + c], wg1[k] = W.wl1[k * MD_GATE + c + D];
  // ---- fusion of the graph rows (BitwiseMultipyLogis) ------------------------------------------------
  {
    float o = 0.f;
#pragma unroll
    for (int k = 0; k < D; ++k) o = fmaf(y[l][k], trc[k], o);
    T[l][c] = tanhf(o + W.bias[c]);
  }
  __syncthreads();
  const float lw = W.logis_w[c], lb = W.logis_b[0];
  for (int k = 0; k < L; ++k) {  // dots[p=l][k] = (T_k . T_l . lw)
    const float d = group_sum64(T[k][c] * T[l][c] * lw, scratch, l, c);
    if (c == 0) dots[l][k] = d;
  }
  __syncthreads();
  float x = T[l][c];
  {
    float a[MD_MAX_LAYERS], mx = -1e30f, den = 0.f;
    for (int k = 0; k < L; ++k) {
      a[k] = 1.f / (1.f + expf(-(dots[l][k] + lb)));
      mx = fmaxf(mx, a[k]);
    }
    for (int k = 0; k < L; ++k) {
      a[k] = expf(a[k] - mx);
      den += a[k];
    }
    for (int k = 0; k < L; ++k)
      if (k != l) x = fmaf(a[k] / den, T[k][c], x);
  }
  {
    const float nrm = sqrtf(group_sum64(x * x, scratch, l, c));
    x = x / fmaxf(nrm, 1e-12f);
  }
  __syncthreads();
  y[l][c] = x;  // yf_l
  __syncthreads();
  // s_l = yf_l . cross_product ; gate_l = relu(yf_l @ w_layer1) @ w_layer2
  {
    const float s = group_sum64(x * W.cross[c], scratch, l, c);
    if (c == 0) sval[l] = s;
  }
  static_assert(MD_GATE == 2 * D, "the gate's hidden layer is two columns per thread");
  {
    float o0 = 0.f, o1 = 0.f;
#pragma unroll
    for (int k = 0; k < D; ++k) {
      o0 = fmaf(y[l][k], wg0[k], o0);
      o1 = fmaf(y[l][k], wg1[k], o1);
    }
    hid[l][c] = fmaxf(o0, 0.f) * W.wl2[c];
    hid[l][c + D] = fmaxf(o1, 0.f) * W.wl2[c + D];
  }
  __syncthreads();
  {
    const float gsum = group_sum64(hid[l][c] + hid[l][c + D], scratch, l, c);
    if (c == 0) gate[l] = gsum;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float* gc = K.gconst + (size_t)g * MD_GCONST;
    float mx = gate[0];
    for (int k = 1; k < L; ++k) mx = fmaxf(mx, gate[k]);
    float e[MD_MAX_LAYERS], den = 0.f;
    for (int k = 0; k < L; ++k) {
      e[k] = expf(gate[k] - mx);
      den += e[k];
    }
    for (int k = 0; k < L; ++k) {
      const float* ax = S.aux + ((size_t)g * L + k) * MD_AUX;
      float aq = 0.f;
      for (int i = 0; i < MD_AUX; ++i) aq = fmaf(ax[i], W.h2[MD_HIDDEN + i], aq);
      gc[k] = sval[k];
      gc[4 + k] = e[k] / den;
      gc[8 + k] = aq;
    }
  }
}

}  // namespace md

// =================================================================================================
// host entry points
// =================================================================================================
bool md_h0_on_the_fly(const md_graphs_t* g);

static int check_common(const md_graphs_t* g, const md_state_t* s, const md_weights_t* w, const md_work_t* k) {
  if (!g || !s || !w || !k) return MD_EINVAL;
  if (g->n_layers < 1 || g->n_layers > MD_MAX_LAYERS) return MD_ELAYERS;
  if (g->stride % MD_TILE_M != 0 || g->n_total != g->n_graphs * g->stride) return MD_EINVAL;
  return 0;
}

extern "C" int md_embed_init(const md_graphs_t* g, const md_state_t* s, const md_weights_t* w, const md_work_t* k,
                             md_stream_t stream) {
  int rc = check_common(g, s, w, k);
  if (rc) return rc;
  if (md_h0_on_the_fly(g)) return 0;  // large unit-cost graphs: iteration 0 works from the alive degrees (md_agg.cu)
  dim3 grid(g->n_graphs * (g->stride / MD_TILE_M), g->n_layers);
  md::embed_init_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*g, *s, *w, *k);
  MD_CHECK_LAUNCH();
  return 0;
}

extern "C" int md_graph_embed(const md_graphs_t* g, const md_state_t* s, const md_weights_t* w, const md_work_t* k,
                              md_stream_t stream) {
  int rc = check_common(g, s, w, k);
  if (rc) return rc;
  if (g->stride / MD_TILE_M > md::PSUM_REDUCE_MIN_TILES) {
    md::psum_reduce_kernel<<<MD_BP_ITERS * g->n_layers * g->n_graphs, 1024, 0, (cudaStream_t)stream>>>(*g, *s, *k);
    MD_CHECK_LAUNCH();
  }
  md::graph_embed_kernel<<<g->n_graphs, g->n_layers * md::D, 0, (cudaStream_t)stream>>>(*g, *s, *w, *k);
  MD_CHECK_LAUNCH();
  return 0;
}

extern "C" int64_t md_psum_floats(int32_t n_layers, int32_t n_total) {
  return (int64_t)MD_BP_ITERS * n_layers * (n_total / MD_TILE_M) * MD_EMBED;
}
