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
// md_embed_iter : persistent, warp-specialised CTAs (one per SM)
//   8 producer warps gather tile t+1 (alive-neighbour sums + own rows) into one half of a double-buffered
//   shared-memory A tile while 8 consumer warps run the dense transform + normalise epilogue on tile t.
//   Hand-off through named barriers (full[2] / empty[2]); the merged 128x64 weight stays resident in smem.
// =================================================================================================
constexpr int IT_PROD = 512;  // 16 gather warps
constexpr int IT_CONS = 256;
constexpr int IT_T = IT_PROD + IT_CONS;
constexpr int EMAX = 2048;  // staged alive-neighbour slots per tile
constexpr int SEG = 32;     // entries per gather segment
constexpr int MAXX = EMAX / SEG;  // most extra segments a staged tile can have
constexpr int GCH = 8;      // neighbour rows in flight per half-warp
struct IterSmem {
  float A[2][TILE_M * A_LD];  // [buf][row][0:64] = neighbour sum P, [row][64:128] = own embedding H
  float Wm[2 * D * D];        // merged weight [128][64]: rows 0..63 = w13, 64..127 = w23
  float red[8 * D];           // partial column sums
  float extra[MAXX * D];      // partial sums of the 2nd.. segments of long (hub) rows
  int eu[EMAX];               // staged alive-neighbour ids of the tile
  int v[2][TILE_M];           // global node id of the row, -1 = padding row
  int off[TILE_M + 1];        // first staged entry of each tile row (from the alive CSR row pointers)
  int seg_beg[TILE_M + MAXX]; // work list: first entry / length / destination of every segment
  int seg_len[TILE_M + MAXX];
  int xbase[TILE_M];          // first `extra` slot of the row
  int mcount, nsegs;
  short seg_dst[TILE_M + MAXX];
  unsigned char seg_row[TILE_M + MAXX];
  unsigned char nseg[TILE_M];
  unsigned char mrow[TILE_M];  // rows with more than one segment
};

enum { BAR_FULL0 = 1, BAR_EMPTY0 = 3, BAR_PROD = 5 };
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

__device__ __forceinline__ void tile_gemm_128(const float* __restrict__ A, const float* __restrict__ Wm, int ty, int tx,
                                              float (&acc)[8][4]) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[i][c] = 0.f;
#pragma unroll 2
  for (int k0 = 0; k0 < 2 * D; k0 += 4) {
    float4 a[8], w[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = *reinterpret_cast<const float4*>(A + (ty + 16 * i) * A_LD + k0);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) w[kk] = *reinterpret_cast<const float4*>(Wm + (k0 + kk) * D + 4 * tx);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      acc[i][0] = fmaf(a[i].x, w[0].x, acc[i][0]);
      acc[i][1] = fmaf(a[i].x, w[0].y, acc[i][1]);
      acc[i][2] = fmaf(a[i].x, w[0].z, acc[i][2]);
      acc[i][3] = fmaf(a[i].x, w[0].w, acc[i][3]);
      acc[i][0] = fmaf(a[i].y, w[1].x, acc[i][0]);
      acc[i][1] = fmaf(a[i].y, w[1].y, acc[i][1]);
      acc[i][2] = fmaf(a[i].y, w[1].z, acc[i][2]);
      acc[i][3] = fmaf(a[i].y, w[1].w, acc[i][3]);
      acc[i][0] = fmaf(a[i].z, w[2].x, acc[i][0]);
      acc[i][1] = fmaf(a[i].z, w[2].y, acc[i][1]);
      acc[i][2] = fmaf(a[i].z, w[2].z, acc[i][2]);
      acc[i][3] = fmaf(a[i].z, w[2].w, acc[i][3]);
      acc[i][0] = fmaf(a[i].w, w[3].x, acc[i][0]);
      acc[i][1] = fmaf(a[i].w, w[3].y, acc[i][1]);
      acc[i][2] = fmaf(a[i].w, w[3].z, acc[i][2]);
      acc[i][3] = fmaf(a[i].w, w[3].w, acc[i][3]);
    }
  }
}

struct TileIter {  // the tile sequence both roles walk in lock step
  int tile, n_tiles, tiles_per_layer, tpg;
  int l, gt, g, t, nact;
  __device__ __forceinline__ bool next(const md_state_t& S) {
    for (; tile < n_tiles; tile += gridDim.x) {
      l = tile / tiles_per_layer;
      gt = tile - l * tiles_per_layer;
      g = gt / tpg;
      t = gt - g * tpg;
      nact = S.n_act[g];
      if (t * TILE_M < nact) return true;
    }
    return false;
  }
};

__global__ void __launch_bounds__(IT_T, 1) embed_iter_kernel(md_graphs_t G, md_state_t S, md_weights_t W, md_work_t K,
                                                            int iter, int dbg) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  IterSmem& sm = *reinterpret_cast<IterSmem*>(smem_raw);
  const int tid = threadIdx.x;
  TileIter ti;
  ti.tpg = G.stride / TILE_M;
  ti.tiles_per_layer = G.n_graphs * ti.tpg;
  ti.n_tiles = ti.tiles_per_layer * G.n_layers;
  ti.tile = blockIdx.x;

  if (tid < IT_PROD) {
    // ===================================== PRODUCER =====================================
    // Streams the alive CSR the env kernel emitted in compact order: the rows of a tile are consecutive, so the
    // tile's neighbour list is one contiguous, coalesced copy and row boundaries come from the row pointers.
    const int lane = tid & 31, warp = tid >> 5;
    int it = 0;
    for (; ti.next(S); ti.tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const int l = ti.l, g = ti.g, t = ti.t;
      const int rows = min(TILE_M, ti.nact - t * TILE_M);
      const int32_t* __restrict__ aptr = S.aptr[l] + g * G.stride + t * TILE_M;
      const int32_t* __restrict__ acol = S.acol[l];
      const float* __restrict__ Hin = K.h[iter & 1][l];
      float* A = sm.A[buf];
      if (it >= 2) bar_sync(BAR_EMPTY0 + buf, IT_T);  // consumers are done with this buffer

      // ---- stage A: row descriptors (one coalesced load each) ---------------------------------------------
      const int b0 = aptr[0];
      if (tid <= TILE_M) {
        const int r = min(tid, rows);
        sm.off[tid] = aptr[r] - b0;
        if (tid < TILE_M) sm.v[buf][tid] = tid < rows ? S.act_list[g * G.stride + t * TILE_M + tid] : -1;
      }
      if (tid == 0) sm.mcount = 0;
      bar_sync(BAR_PROD, IT_PROD);
      const int etot = sm.off[TILE_M];
      const bool staged = etot <= EMAX;
      // ---- stage B: the tile's alive-neighbour ids (contiguous in acol) + the balanced work list ---------------
      // A row's neighbour list is cut into segments of <= SEG entries; segments are dealt round-robin to the 32
      // half-warps, so a hub row (degree in the hundreds on the power-law GMM graphs) is summed by many half-warps
      // in parallel instead of serialising one.  Segment 0 of a row lands in the A tile, the others in `extra`
      // and are folded in segment order afterwards: the f32 result stays deterministic.
      if (staged)
        for (int e = tid; e < etot; e += IT_PROD) sm.eu[e] = acol[b0 + e];
      if (warp == 0) {
        int nx[4], tot = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int n = sm.off[4 * lane + q + 1] - sm.off[4 * lane + q];
          nx[q] = staged ? max(0, (n - 1) / SEG) : 0;  // extra segments of the row
          tot += nx[q];
        }
        int incl = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int n = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += n;
        }
        int xb = incl - tot;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int r = 4 * lane + q;
          const int b = sm.off[r], n = sm.off[r + 1] - b;
          sm.xbase[r] = xb;
          sm.nseg[r] = (unsigned char)(1 + nx[q]);
          if (nx[q] > 0) sm.mrow[atomicAdd(&sm.mcount, 1)] = (unsigned char)r;
          for (int k = 0; k <= nx[q]; ++k) {
            const int sidx = r + xb + k;
            sm.seg_row[sidx] = (unsigned char)r;
            sm.seg_beg[sidx] = b + k * SEG;
            sm.seg_len[sidx] = staged ? min(SEG, n - k * SEG) : n;
            sm.seg_dst[sidx] = (short)(k == 0 ? -1 : xb + k - 1);
          }
          xb += nx[q];
        }
        if (lane == 31) sm.nsegs = TILE_M + incl;
      }
      bar_sync(BAR_PROD, IT_PROD);

      // ---- stage C: gather.  A half-warp sums one segment; a lane holds 4 consecutive columns, so one LDG.128
      //      per half-warp fetches a whole 256-byte neighbour row -----------------------------------------------
      if (!(dbg & 1)) {
        const int half = lane >> 4, l16 = lane & 15;
        const float4* __restrict__ H4 = reinterpret_cast<const float4*>(Hin);
        const int nsegs = sm.nsegs;
        for (int s0 = warp * 2; s0 < nsegs; s0 += IT_PROD / 16) {
          const int sidx = s0 + half;
          const bool valid = sidx < nsegs;
          const int b = valid ? sm.seg_beg[sidx] : 0, n = valid ? sm.seg_len[sidx] : 0;
          const int nmax = max(n, __shfl_xor_sync(0xffffffffu, n, 16));
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int k0 = 0; k0 < nmax; k0 += GCH) {
            float4 val[GCH];
#pragma unroll
            for (int j = 0; j < GCH; ++j) {
              val[j] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (k0 + j < n) {
                const int u = staged ? sm.eu[b + k0 + j] : acol[b0 + b + k0 + j];
                val[j] = __ldg(H4 + (size_t)u * (D / 4) + l16);
              }
            }
#pragma unroll
            for (int j = 0; j < GCH; ++j) {  // in CSR order
              if (k0 + j < n) {
                acc.x += val[j].x;
                acc.y += val[j].y;
                acc.z += val[j].z;
                acc.w += val[j].w;
              }
            }
          }
          if (valid) {
            const int dst = sm.seg_dst[sidx];
            float* out = dst < 0 ? A + sm.seg_row[sidx] * A_LD : sm.extra + dst * D;
            *reinterpret_cast<float4*>(out + 4 * l16) = acc;
          }
        }
        // own rows -> A[:, 64:128]
        float4 own[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int v = sm.v[buf][warp * 8 + 2 * i + half];
          own[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (v >= 0) own[i] = __ldg(H4 + (size_t)v * (D / 4) + l16);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
          *reinterpret_cast<float4*>(A + (warp * 8 + 2 * i + half) * A_LD + D + 4 * l16) = own[i];
        // fold the extra segments of hub rows, in segment order
        const int mcount = sm.mcount;
        if (mcount > 0) {
          bar_sync(BAR_PROD, IT_PROD);
          for (int m = warp * 2 + half; m < mcount; m += IT_PROD / 16) {
            const int r = sm.mrow[m];
            float4 acc = *reinterpret_cast<const float4*>(A + r * A_LD + 4 * l16);
            const int xb = sm.xbase[r], nx = sm.nseg[r] - 1;
            for (int k = 0; k < nx; ++k) {
              const float4 x = *reinterpret_cast<const float4*>(sm.extra + (xb + k) * D + 4 * l16);
              acc.x += x.x;
              acc.y += x.y;
              acc.z += x.z;
              acc.w += x.w;
            }
            *reinterpret_cast<float4*>(A + r * A_LD + 4 * l16) = acc;
          }
        }
      }
      __threadfence_block();
      bar_arrive(BAR_FULL0 + buf, IT_T);  // hand the tile to the consumers
      bar_sync(BAR_PROD, IT_PROD);        // all producer warps finished writing A (needed for the column sums)

      // ---- per-tile partial of the subgraph sum (old embeddings), fixed order => deterministic --------------
      {
        const int c = tid & 63, part = tid >> 6;  // 8 row eighths x 64 columns
        float s = 0.f;
        for (int r = part * 16; r < part * 16 + 16; ++r) s += A[r * A_LD + D + c];
        sm.red[part * D + c] = s;
      }
      bar_sync(BAR_PROD, IT_PROD);
      if (tid < D) {
        float s = sm.red[tid];
#pragma unroll
        for (int q = 1; q < 8; ++q) s += sm.red[q * D + tid];
        K.psum[(((size_t)iter * G.n_layers + l) * ti.tiles_per_layer + ti.gt) * D + tid] = s;
      }
    }
    // drain: match the consumers' last empty-arrivals so no barrier is left half-arrived
    for (int j = max(0, it - 2); j < it; ++j) bar_sync(BAR_EMPTY0 + (j & 1), IT_T);
  } else {
    // ===================================== CONSUMER =====================================
    const int ct = tid - IT_PROD;
    const int tx = ct & 15, ty = ct >> 4;
    // merged weight, resident for the lifetime of the CTA (consumers only touch Wm)
    for (int i = ct; i < D * D / 4; i += IT_CONS) {
      reinterpret_cast<float4*>(sm.Wm)[i] = reinterpret_cast<const float4*>(W.w13)[i];
      reinterpret_cast<float4*>(sm.Wm + D * D)[i] = reinterpret_cast<const float4*>(W.w23)[i];
    }
    bar_sync(BAR_PROD + 1, IT_CONS);
    int it = 0;
    for (; ti.next(S); ti.tile += gridDim.x, ++it) {
      const int buf = it & 1;
      float* __restrict__ Hout = K.h[(iter + 1) & 1][ti.l];
      bar_sync(BAR_FULL0 + buf, IT_T);
      float acc[8][4];
      if (dbg & 2) {
        __threadfence_block();
        bar_arrive(BAR_EMPTY0 + buf, IT_T);
        continue;
      }
      tile_gemm_128(sm.A[buf], sm.Wm, ty, tx, acc);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float ss = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          acc[i][c] = fmaxf(acc[i][c], 0.f);
          ss = fmaf(acc[i][c], acc[i][c], ss);
        }
        ss += __shfl_xor_sync(0xffffffffu, ss, 1);
        ss += __shfl_xor_sync(0xffffffffu, ss, 2);
        ss += __shfl_xor_sync(0xffffffffu, ss, 4);
        ss += __shfl_xor_sync(0xffffffffu, ss, 8);
        // x / max(|x|, eps) as x * rcp_rn(.): one correctly rounded reciprocal per row instead of 4 IEEE divisions
        // per thread (the division subroutine was 30% of the first version's instructions); differs by <= 1 ulp
        const float inv = __frcp_rn(fmaxf(sqrtf(ss), 1e-12f));
        const int v = sm.v[buf][ty + 16 * i];
        if (v >= 0)
          *reinterpret_cast<float4*>(Hout + (size_t)v * D + 4 * tx) =
              make_float4(acc[i][0] * inv, acc[i][1] * inv, acc[i][2] * inv, acc[i][3] * inv);
      }
      __threadfence_block();
      bar_arrive(BAR_EMPTY0 + buf, IT_T);
    }
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
  for (int it = 0; it < MD_BP_ITERS; ++it) {
    float s = 0.f;
    const float* ps = K.psum + (((size_t)it * L + l) * tiles_per_layer + (size_t)g * tpg) * D + c;
    for (int t = 0; t < ntile; ++t) s += ps[(size_t)t * D];
    sbuf[l][c] = s;
    __syncthreads();
    float o = 0.f;
    for (int k = 0; k < D; ++k) o = fmaf(sbuf[l][k], W.w13[k * D + c], o);
    for (int k = 0; k < D; ++k) o = fmaf(y[l][k], W.w23[k * D + c], o);
    o = fmaxf(o, 0.f);
    const float nrm = sqrtf(group_sum64(o * o, scratch, l, c));  // contains __syncthreads: y reads are done
    y[l][c] = o / fmaxf(nrm, 1e-12f);
    __syncthreads();
  }

  // ---- fusion of the graph rows (BitwiseMultipyLogis) ------------------------------------------------
  {
    float o = 0.f;
    for (int k = 0; k < D; ++k) o = fmaf(y[l][k], W.trans[k * D + c], o);
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
  for (int j = c; j < MD_GATE; j += D) {
    float o = 0.f;
    for (int k = 0; k < D; ++k) o = fmaf(y[l][k], W.wl1[k * MD_GATE + j], o);
    hid[l][j] = fmaxf(o, 0.f) * W.wl2[j];
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
int g_md_debug_flags = 0;
extern "C" void md_debug_set_flags(int32_t f) { g_md_debug_flags = f; }

static int check_common(const md_graphs_t* g, const md_state_t* s, const md_weights_t* w, const md_work_t* k) {
  if (!g || !s || !w || !k) return MD_EINVAL;
  if (g->n_layers < 1 || g->n_layers > MD_MAX_LAYERS) return MD_ELAYERS;
  if (g->stride % MD_TILE_M != 0 || g->n_total != g->n_graphs * g->stride) return MD_EINVAL;
  return 0;
}

static int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

extern "C" int md_embed_init(const md_graphs_t* g, const md_state_t* s, const md_weights_t* w, const md_work_t* k,
                             md_stream_t stream) {
  int rc = check_common(g, s, w, k);
  if (rc) return rc;
  dim3 grid(g->n_graphs * (g->stride / MD_TILE_M), g->n_layers);
  md::embed_init_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*g, *s, *w, *k);
  MD_CHECK_LAUNCH();
  return 0;
}

extern "C" int md_embed_iter(const md_graphs_t* g, const md_state_t* s, const md_weights_t* w, const md_work_t* k,
                             int32_t iter, md_stream_t stream) {
  int rc = check_common(g, s, w, k);
  if (rc) return rc;
  if (iter < 0 || iter >= MD_BP_ITERS) return MD_EINVAL;
  static bool configured = false;
  const int smem = (int)sizeof(md::IterSmem);
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(md::embed_iter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  const int n_tiles = g->n_graphs * (g->stride / MD_TILE_M) * g->n_layers;
  int grid = sm_count();
  if (grid > n_tiles) grid = n_tiles;
  md::embed_iter_kernel<<<grid, md::IT_T, smem, (cudaStream_t)stream>>>(*g, *s, *w, *k, iter, g_md_debug_flags);
  MD_CHECK_LAUNCH();
  return 0;
}

extern "C" int md_graph_embed(const md_graphs_t* g, const md_state_t* s, const md_weights_t* w, const md_work_t* k,
                              md_stream_t stream) {
  int rc = check_common(g, s, w, k);
  if (rc) return rc;
  md::graph_embed_kernel<<<g->n_graphs, g->n_layers * md::D, 0, (cudaStream_t)stream>>>(*g, *s, *w, *k);
  MD_CHECK_LAUNCH();
  return 0;
}

extern "C" int64_t md_psum_floats(int32_t n_layers, int32_t n_total) {
  return (int64_t)MD_BP_ITERS * n_layers * (n_total / MD_TILE_M) * MD_EMBED;
}
