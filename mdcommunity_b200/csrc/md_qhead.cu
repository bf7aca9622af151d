// Inter-layer fusion of the node rows + Q head + masked arg-max, one pass over the active nodes.
//
// Replaces (reference, /root/reference/code/MultiDismantler_unit_cost):
//   MRGNN/mutil_layer_weight.py:266-285,294-301,304-313  BitwiseMultipyLogis / layer_bitwise / LogisticVector
//   MultiDismantler_net_graphsage.py:323-393               fusion call, normalise, Q head, layer gate, mix
//   MultiDismantler_torch.py:286-300, :769, :799-807       scatter to [N] with -inf for inactive nodes + arg-max
//
// Per node v of graph g (L layers, H_l = embedding after 3 iterations):
//   T_l = tanh(H_l @ trans + bias)
//   X_p = T_p + sum_{k != p} softmax_k( sigmoid((T_k * T_p) . lw + lb) ) T_k ;  Hf_p = X_p / max(|X_p|, eps)
//   q_p = relu((Hf_p * s_p[g]) @ h1) . h2[:32] + auxq_p[g] ;   q = sum_p gate_p[g] q_p
// s_p, gate_p, auxq_p only depend on the graph row and come from md_graph_embed (gconst).
// The [n,64,64] outer-product temporary of net_graphsage.py:356-363 is never formed: e = Hf * (y . cross).
// Inactive / removed nodes are simply not in act_list, which is the -inf mask of Predict.
#include "md_common.cuh"

namespace md {

constexpr int QH_T = 256;

template <int L>
struct QSmem {
  static constexpr int LD = L * D + 4;
  float A[TILE_M * LD];
  float trans[D * D];
  float h1[D * MD_HIDDEN];
  float bias[D];
  float lw[D];
  float h2[MD_HIDDEN];
  int v[TILE_M];
};

template <int L>
__global__ void __launch_bounds__(QH_T, (L == 2 ? 2 : 1)) qhead_kernel(md_graphs_t G, md_state_t S, md_weights_t W,
                                                                      md_work_t K) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using SM = QSmem<L>;
  SM& sm = *reinterpret_cast<SM*>(smem_raw);
  constexpr int LD = SM::LD;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tx = tid & 15, ty = tid >> 4;
  const int tpg = G.stride / TILE_M;
  const int n_tiles = G.n_graphs * tpg;

  for (int i = tid; i < D * D / 4; i += QH_T)
    reinterpret_cast<float4*>(sm.trans)[i] = reinterpret_cast<const float4*>(W.trans)[i];
  for (int i = tid; i < D * MD_HIDDEN / 4; i += QH_T)
    reinterpret_cast<float4*>(sm.h1)[i] = reinterpret_cast<const float4*>(W.h1)[i];
  if (tid < D) {
    sm.bias[tid] = W.bias[tid];
    sm.lw[tid] = W.logis_w[tid];
  }
  if (tid < MD_HIDDEN) sm.h2[tid] = W.h2[tid];
  const float lb = W.logis_b[0];

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int g = tile / tpg, t = tile - g * tpg;
    const int nact = S.n_act[g];
    if (t * TILE_M >= nact) continue;
    const int rows = min(TILE_M, nact - t * TILE_M);
    __syncthreads();
    if (tid < TILE_M) sm.v[tid] = tid < rows ? S.act_list[g * G.stride + t * TILE_M + tid] : -1;
    __syncthreads();
    // load H_l rows: warp w owns rows [16w, 16w+16).  All 16 L row fetches of the warp are issued before the first
    // store (one L2 round trip per tile instead of 16 L / 4: ncu showed 43 % of the kernel's stall samples here)
    const int hsel = MD_BP_ITERS & 1;
    {
      float2 hreg[L][16];
#pragma unroll
      for (int l = 0; l < L; ++l) {
        const float* __restrict__ H = K.h[hsel][l];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int v = sm.v[warp * 16 + i];
          hreg[l][i] = make_float2(0.f, 0.f);
          if (v >= 0) hreg[l][i] = __ldg(reinterpret_cast<const float2*>(H + (size_t)v * D) + lane);
        }
      }
#pragma unroll
      for (int l = 0; l < L; ++l)
#pragma unroll
        for (int i = 0; i < 16; ++i)
          *reinterpret_cast<float2*>(sm.A + (warp * 16 + i) * LD + l * D + 2 * lane) = hreg[l][i];
    }
    __syncthreads();

    // ---- T_l = tanh(H_l @ trans + bias), written back in place (rows are private to a half-warp) -------
    const float4 bv = *reinterpret_cast<const float4*>(sm.bias + 4 * tx);
#pragma unroll
    for (int l = 0; l < L; ++l) {
      float acc[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[i][c] = 0.f;
#pragma unroll 2
      for (int k0 = 0; k0 < D; k0 += 4) {
        float4 a[8], w[4];
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = *reinterpret_cast<const float4*>(sm.A + (ty + 16 * i) * LD + l * D + k0);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) w[kk] = *reinterpret_cast<const float4*>(sm.trans + (k0 + kk) * D + 4 * tx);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
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
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; ++i)
        *reinterpret_cast<float4*>(sm.A + (ty + 16 * i) * LD + l * D + 4 * tx) =
            make_float4(tanhf(acc[i][0] + bv.x), tanhf(acc[i][1] + bv.y), tanhf(acc[i][2] + bv.z),
                        tanhf(acc[i][3] + bv.w));
    }
    __syncwarp();

    // ---- fusion weights, normalise, scale by s_p[g]; e_p overwrites T_p in place -----------------------
    float sg[L], gw[L], aq[L];
#pragma unroll
    for (int p = 0; p < L; ++p) {
      sg[p] = K.gconst[(size_t)g * MD_GCONST + p];
      gw[p] = K.gconst[(size_t)g * MD_GCONST + 4 + p];
      aq[p] = K.gconst[(size_t)g * MD_GCONST + 8 + p];
    }
    const float4 lwv = *reinterpret_cast<const float4*>(sm.lw + 4 * tx);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float* rowp = sm.A + (ty + 16 * i) * LD + 4 * tx;
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
          d += __shfl_xor_sync(0xffffffffu, d, 1);
          d += __shfl_xor_sync(0xffffffffu, d, 2);
          d += __shfl_xor_sync(0xffffffffu, d, 4);
          d += __shfl_xor_sync(0xffffffffu, d, 8);
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
        ss += __shfl_xor_sync(0xffffffffu, ss, 1);
        ss += __shfl_xor_sync(0xffffffffu, ss, 2);
        ss += __shfl_xor_sync(0xffffffffu, ss, 4);
        ss += __shfl_xor_sync(0xffffffffu, ss, 8);
        const float inv = __frcp_rn(fmaxf(sqrtf(ss), 1e-12f));
        const float s = sg[p];
        *reinterpret_cast<float4*>(rowp + p * D) =
            make_float4(x.x * inv * s, x.y * inv * s, x.z * inv * s, x.w * inv * s);
      }
    }
    __syncwarp();

    // ---- q_p = relu(e_p @ h1) . h2[:32] + auxq_p ;  q = sum_p gate_p q_p -------------------------------
    float qrow[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) qrow[i] = 0.f;
    const float2 h2v = *reinterpret_cast<const float2*>(sm.h2 + 2 * tx);
#pragma unroll
    for (int p = 0; p < L; ++p) {
      float acc[8][2];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = 0.f;
#pragma unroll 2
      for (int k0 = 0; k0 < D; k0 += 4) {
        float4 a[8];
        float2 w[4];
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = *reinterpret_cast<const float4*>(sm.A + (ty + 16 * i) * LD + p * D + k0);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) w[kk] = *reinterpret_cast<const float2*>(sm.h1 + (k0 + kk) * MD_HIDDEN + 2 * tx);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float av[4] = {a[i].x, a[i].y, a[i].z, a[i].w};
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            acc[i][0] = fmaf(av[kk], w[kk].x, acc[i][0]);
            acc[i][1] = fmaf(av[kk], w[kk].y, acc[i][1]);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float hq = fmaxf(acc[i][0], 0.f) * h2v.x;
        hq = fmaf(fmaxf(acc[i][1], 0.f), h2v.y, hq);
        hq += __shfl_xor_sync(0xffffffffu, hq, 1);
        hq += __shfl_xor_sync(0xffffffffu, hq, 2);
        hq += __shfl_xor_sync(0xffffffffu, hq, 4);
        hq += __shfl_xor_sync(0xffffffffu, hq, 8);
        qrow[i] = fmaf(gw[p], hq + aq[p], qrow[i]);
      }
    }

    // ---- write q, masked arg-max: highest q, ties -> lowest node id ------------------------------------
    unsigned long long key = 0ull;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int v = sm.v[ty + 16 * i];
      if (v >= 0 && tx == 0) {
        K.q[v] = qrow[i];
        const uint32_t local = (uint32_t)(v - g * G.stride);
        const unsigned long long k2 = ((unsigned long long)orderable(qrow[i]) << 32) | (0xffffffffu - local);
        key = k2 > key ? k2 : key;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
      key = other > key ? other : key;
    }
    if (lane == 0 && key != 0ull) atomicMax(&S.best[g], key);
  }
}

// Top-k selection for GetSolution's stepRatio mode (U/MultiDismantler_torch.py:711-736, np.argsort(-pred)[:step]):
// one CTA per graph ranks its active nodes by (q desc, id asc) and writes the first k local ids.
__global__ void __launch_bounds__(256) topk_kernel(md_graphs_t G, md_state_t S, md_work_t K, int32_t* out, int k) {
  const int g = blockIdx.x;
  const int nact = S.done[g] ? 0 : S.n_act[g];
  const int base = g * G.stride;
  for (int i = threadIdx.x; i < k; i += blockDim.x) out[(size_t)g * k + i] = -1;
  __syncthreads();
  for (int i = threadIdx.x; i < nact; i += blockDim.x) {
    const int v = S.act_list[base + i];
    const float qv = K.q[v];
    int rank = 0;
    for (int j = 0; j < nact; ++j) {
      const int u = S.act_list[base + j];
      const float qu = K.q[u];
      rank += (qu > qv) || (qu == qv && u < v);
    }
    if (rank < k) out[(size_t)g * k + rank] = v - base;
  }
}

}  // namespace md

int md_qhead_tc_launch(const md_graphs_t* g, const md_state_t* s, const md_weights_t* w, const md_work_t* k, cudaStream_t st,
                       int* launched);

extern "C" int md_qhead(const md_graphs_t* g, const md_state_t* s, const md_weights_t* w, const md_work_t* k,
                        md_stream_t stream) {
  if (!g || !s || !w || !k) return MD_EINVAL;
  if (g->stride % MD_TILE_M != 0 || g->n_total != g->n_graphs * g->stride) return MD_EINVAL;
  {  // tensor-core kernel (md_qhead_tc.cu) unless md_debug_set_flags bit 3 asks for the FP32-FMA kernel below
    int launched = 0;
    const int rc = md_qhead_tc_launch(g, s, w, k, (cudaStream_t)stream, &launched);
    if (rc) return rc;
    if (launched) return 0;
  }
  int dev = 0, nsm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  const int n_tiles = g->n_graphs * (g->stride / MD_TILE_M);
  cudaStream_t st = (cudaStream_t)stream;
  if (g->n_layers == 2) {
    static md::PerDevice<bool> configured_pd;
  bool& configured = configured_pd.get();
    const int smem = (int)sizeof(md::QSmem<2>);
    if (!configured) {
      cudaError_t e = cudaFuncSetAttribute(md::qhead_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      if (e != cudaSuccess) return (int)e;
      configured = true;
    }
    int grid = 2 * nsm < n_tiles ? 2 * nsm : n_tiles;
    md::qhead_kernel<2><<<grid, md::QH_T, smem, st>>>(*g, *s, *w, *k);
  } else if (g->n_layers == 3) {
    static md::PerDevice<bool> configured_pd;
  bool& configured = configured_pd.get();
    const int smem = (int)sizeof(md::QSmem<3>);
    if (!configured) {
      cudaError_t e = cudaFuncSetAttribute(md::qhead_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      if (e != cudaSuccess) return (int)e;
      configured = true;
    }
    int grid = nsm < n_tiles ? nsm : n_tiles;
    md::qhead_kernel<3><<<grid, md::QH_T, smem, st>>>(*g, *s, *w, *k);
  } else {
    return MD_ELAYERS;
  }
  MD_CHECK_LAUNCH();
  return 0;
}

extern "C" int md_select_topk(const md_graphs_t* g, const md_state_t* s, const md_work_t* k, int32_t* out_local_ids,
                              int32_t topk, md_stream_t stream) {
  if (!g || !s || !k || !out_local_ids || topk < 1) return MD_EINVAL;
  md::topk_kernel<<<g->n_graphs, 256, 0, (cudaStream_t)stream>>>(*g, *s, *k, out_local_ids, topk);
  MD_CHECK_LAUNCH();
  return 0;
}
