// md_embed_iter: one message-passing iteration for every layer of the batch, fused:
//   neighbour sum over the alive edges (torch_sparse.spmm n2n), dense transforms, relu, L2 normalise and the per-tile
//   column sums of the old embeddings (torch_sparse.spmm subg).
// Replaces /root/reference/code/MultiDismantler_unit_cost/MultiDismantler_net_graphsage.py:288-310 (:297-298 subg).
//
// Persistent, warp-specialised CTA per SM (24 warps), 128-row tiles, double-buffered shared A tile [P | H]:
//   warps 0..15  gather warps.  Measured on B200 (scripts/microbench/gather_lat.cu): one warp sustains only ~1.6 B/cycle of
//                256-byte row gathers however many loads it has in flight, an L2 hit costs ~420 ns, the SM tops out near
//                37 B/cycle with 32 warps.  So gather parallelism has to come from warps that never wait for each
//                other: every warp owns 8 consecutive rows of the tile, fetches their slice of the alive CSR itself
//                (prefetched one tile ahead into a warp-private staging area) and sums the short rows (<= LONG
//                neighbours) without any barrier.  Long (hub) rows are appended to a per-tile list and, after one
//                gather-wide barrier, cut into LONG-entry segments dealt round-robin to all 32 half-warps; partial sums
//                are folded in segment order (deterministic).
//   warps 16..23 consumers: column sums of H, [P|H] @ [W13;W23] (FP32 FFMA, 8x4 register micro-tiles, weights resident
//                in shared memory), relu, row normalise, store.
//   Hand-off through named barriers: full[2] (gather -> consumer), empty[2] (consumer -> gather).
#include "md_common.cuh"

namespace md {

constexpr int IT_GATH = 512;  // 16 gather warps
constexpr int IT_CONS = 256;  // 8 consumer warps
constexpr int IT_T = IT_GATH + IT_CONS;
constexpr int NGW = IT_GATH / 32;
constexpr int LONG = 16;      // rows with more alive neighbours are handled cooperatively
constexpr int WSTG = 8 * LONG;  // warp-private staging: neighbour ids of 8 short rows
constexpr int MAXLONG = TILE_M;  // every row of a tile may be long
constexpr int MAXXS = 96;     // extra (2nd..) segments per tile; more -> whole-row fallback
constexpr int NACT_CACHE = 1024;

struct IterSmem {
  float A[2][TILE_M * A_LD];  // [buf][row][0:64] = neighbour sum P, [row][64:128] = own embedding H
  float Wm[2 * D * D];        // merged weight [128][64]: rows 0..63 = w13, 64..127 = w23
  float extra[MAXXS * D];     // partial sums of the 2nd.. segments of long rows
  float red[2][4 * D];        // partial column sums (consumers)
  int wstg[2][NGW][WSTG];     // [parity][warp] prefetched neighbour ids of the warp's rows
  int wptr[2][NGW][12];       // [parity][warp] row pointers (9) of the warp's rows, relative to the first; [9] = base
  int wv[2][NGW][8];          // [parity][warp] global node ids of the warp's rows
  int v[2][TILE_M];           // global node id of the row, -1 = padding row
  int nact[NACT_CACHE];       // n_act per graph (avoids a dependent global load per tile)
  int long_row[MAXLONG];      // tile rows with > LONG alive neighbours
  int long_beg[MAXLONG];      // absolute first entry in acol
  int long_n[MAXLONG];
  int long_x[MAXLONG];        // first extra slot
  int nlong, nxtra;
};

enum { BAR_FULL0 = 1, BAR_EMPTY0 = 3, BAR_GATH = 5, BAR_CONS = 6 };
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

struct TileIter {  // the tile sequence all roles walk in lock step
  int tile, n_tiles, tiles_per_layer, tpg;
  int l, gt, g, t, nact;
  const int* nact_tab;  // shared-memory copy of n_act (or the global array when the batch is too large)
  __device__ __forceinline__ bool next() {
    for (; tile < n_tiles; tile += gridDim.x) {
      l = tile / tiles_per_layer;
      gt = tile - l * tiles_per_layer;
      g = gt / tpg;
      t = gt - g * tpg;
      nact = nact_tab[g];
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
  const int lane = tid & 31, warp = tid >> 5;
  const bool cached = G.n_graphs <= NACT_CACHE;
  if (cached)
    for (int i = tid; i < G.n_graphs; i += IT_T) sm.nact[i] = S.n_act[i];
  if (tid == 0) sm.nlong = 0, sm.nxtra = 0;
  __syncthreads();
  TileIter ti;
  ti.tpg = G.stride / TILE_M;
  ti.tiles_per_layer = G.n_graphs * ti.tpg;
  ti.n_tiles = ti.tiles_per_layer * G.n_layers;
  ti.tile = blockIdx.x;
  ti.nact_tab = cached ? sm.nact : S.n_act;

  if (tid < IT_GATH) {
    // ===================================== GATHER WARPS =====================================
    const int half = lane >> 4, l16 = lane & 15;
    const int hw = warp * 2 + half;
    const int r0 = warp * 8;  // this warp's rows of every tile

    // descriptor prefetch of a tile into staging slot `par`: row pointers, node ids, neighbour ids of short rows
    const bool pf_l2 = !(dbg & 64) && (size_t)G.n_total * (D * 4) * G.n_layers > (size_t)(96u << 20);
    auto prefetch = [&](const TileIter& tt, int par) {
      const int rows = min(TILE_M, tt.nact - tt.t * TILE_M);
      const int32_t* __restrict__ aptr = S.aptr[tt.l] + tt.g * G.stride + tt.t * TILE_M;
      int ap = 0;
      if (lane < 9) ap = aptr[min(r0 + lane, rows)];
      const int b = __shfl_sync(0xffffffffu, ap, 0);
      if (lane < 9) sm.wptr[par][warp][lane] = ap - b;
      if (lane == 9) sm.wptr[par][warp][9] = b;
      // neighbour ids of the rows that are short; a long row's ids are fetched cooperatively later
      const int32_t* __restrict__ acol = S.acol[tt.l];
      const int nxt = __shfl_down_sync(0xffffffffu, ap, 1);
      int n = (lane < 8) ? nxt - ap : 0;
      if (n > LONG) n = 0;
      // exclusive scan of the short-row lengths over lanes 0..7 -> packed position in the staging area
      int incl = n;
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
        const int x = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += x;
      }
      const int excl = incl - n;
      if (lane < 8) sm.wv[par][warp][lane] = (r0 + lane < rows) ? S.act_list[tt.g * G.stride + tt.t * TILE_M + r0 + lane] : -1;
      // every row r copies its n ids: lanes cooperate, 8 rows x up to 32 ids
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int nr = __shfl_sync(0xffffffffu, n, r);
        const int src0 = __shfl_sync(0xffffffffu, ap, r);
        const int dst0 = __shfl_sync(0xffffffffu, excl, r);
        if (lane < nr) {
          const int u = acol[src0 + lane];
          sm.wstg[par][warp][dst0 + lane] = u;
          if (pf_l2) {  // embeddings larger than L2: start the DRAM -> L2 fetch of the neighbour row one tile ahead
            const char* hp = reinterpret_cast<const char*>(K.h[iter & 1][tt.l]) + (size_t)u * (D * 4);
            asm volatile("prefetch.global.L2 [%0];" ::"l"(hp));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(hp + 128));
          }
        }
      }
    };

    int it = 0;
    bool have = ti.next();
    if (have) prefetch(ti, 0);
    while (have) {
      const int buf = it & 1;
      const TileIter cur = ti;
      // advance to the next tile and prefetch its descriptors (other parity) before touching this tile's rows
      ti.tile += gridDim.x;
      have = ti.next();
      if (have) prefetch(ti, buf ^ 1);
      __syncwarp();

      const int l = cur.l, g = cur.g, t = cur.t;
      const int rows = min(TILE_M, cur.nact - t * TILE_M);
      const int32_t* __restrict__ acol = S.acol[l];
      const float4* __restrict__ H4 = reinterpret_cast<const float4*>(K.h[iter & 1][l]);
      float* A = sm.A[buf];
      if (it >= 2) bar_sync(BAR_EMPTY0 + buf, IT_T);  // consumers are done with this A buffer (and v[buf])

      const int* wp = sm.wptr[buf][warp];
      const int* ws = sm.wstg[buf][warp];
      const int base_e = wp[9];
      if (lane < 8) sm.v[buf][r0 + lane] = sm.wv[buf][warp][lane];
      __syncwarp();
      // ---- short rows: 4 row pairs, one row per half-warp, no barrier --------------------------------------
      int spos = 0;  // running position in the packed staging area (same in both halves: recomputed per pair)
      if (!(dbg & 1)) {
#pragma unroll 1
        for (int i = 0; i < 4; ++i) {
          const int ra = 2 * i, rb = 2 * i + 1;
          int na = wp[ra + 1] - wp[ra], nb = wp[rb + 1] - wp[rb];
          const bool la = na > LONG, lb = nb > LONG;
          if (la && lane == 0) {
            const int k = atomicAdd(&sm.nlong, 1);
            sm.long_row[k] = r0 + ra;
            sm.long_beg[k] = base_e + wp[ra];
            sm.long_n[k] = na;
            sm.long_x[k] = atomicAdd(&sm.nxtra, (na - 1) / LONG);
          }
          if (lb && lane == 0) {
            const int k = atomicAdd(&sm.nlong, 1);
            sm.long_row[k] = r0 + rb;
            sm.long_beg[k] = base_e + wp[rb];
            sm.long_n[k] = nb;
            sm.long_x[k] = atomicAdd(&sm.nxtra, (nb - 1) / LONG);
          }
          if (la) na = 0;
          if (lb) nb = 0;
          const int pa = spos, pb = spos + na;
          spos += na + nb;
          const int n = half ? nb : na;
          const int p = half ? pb : pa;
          const int nmax = max(na, nb);
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int k0 = 0; k0 < nmax; k0 += 8) {
            float4 val[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              val[j] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (k0 + j < n) val[j] = __ldg(H4 + (size_t)ws[p + k0 + j] * (D / 4) + l16);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {  // in CSR order
              if (k0 + j < n) {
                acc.x += val[j].x;
                acc.y += val[j].y;
                acc.z += val[j].z;
                acc.w += val[j].w;
              }
            }
          }
          const bool is_long = half ? lb : la;
          if (!is_long) *reinterpret_cast<float4*>(A + (r0 + 2 * i + half) * A_LD + 4 * l16) = acc;
        }
        // own rows -> A[:, 64:128]
        float4 own[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int v = sm.v[buf][r0 + 2 * i + half];
          own[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (v >= 0) own[i] = __ldg(H4 + (size_t)v * (D / 4) + l16);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
          *reinterpret_cast<float4*>(A + (r0 + 2 * i + half) * A_LD + D + 4 * l16) = own[i];
      }
      // ---- long (hub) rows: cooperative, only when the tile has any ---------------------------------------------
      bar_sync(BAR_GATH, IT_GATH);
      const int nlong = sm.nlong;
      const bool whole_rows = sm.nxtra > MAXXS;  // too many segments for the scratch area: one half-warp per long row
      if (nlong > 0 && !(dbg & 1)) {
        // segment s of long row k: entries [s*LONG, min(n, (s+1)*LONG)); flat segment index over all long rows
        int sidx = hw;
        int k = 0, sbase = 0;  // walk the (short) long-row list
        for (;;) {
          int sseg = 0, n = 0, beg = 0, nrep = 1;
          bool valid;
          if (whole_rows) {
            k = sidx;
            valid = k < nlong;
            if (valid) nrep = (sm.long_n[k] + LONG - 1) / LONG;
          } else {
            while (k < nlong && sidx >= sbase + (sm.long_n[k] + LONG - 1) / LONG) {
              sbase += (sm.long_n[k] + LONG - 1) / LONG;
              ++k;
            }
            valid = k < nlong;
            if (valid) sseg = sidx - sbase;
          }
          if (!__any_sync(0xffffffffu, valid)) break;
          const int nrep_max = max(nrep, __shfl_xor_sync(0xffffffffu, nrep, 16));
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int rep = 0; rep < nrep_max; ++rep) {
            const int sg = whole_rows ? rep : sseg;
            n = 0;
            if (valid && rep < nrep) {
              beg = sm.long_beg[k] + sg * LONG;
              n = min(LONG, sm.long_n[k] - sg * LONG);
            }
            // the half-warp fetches its <= 16 neighbour ids (one per lane), then the rows, 8 in flight
            int id0 = 0;
            if (l16 < n) id0 = acol[beg + l16];
            const int nmax = max(n, __shfl_xor_sync(0xffffffffu, n, 16));
            for (int k0 = 0; k0 < nmax; k0 += 8) {
              float4 val[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const int e = k0 + j;
                const int u = __shfl_sync(0xffffffffu, id0, (half << 4) | (e & 15));
                val[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (e < n) val[j] = __ldg(H4 + (size_t)u * (D / 4) + l16);
              }
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                if (k0 + j < n) {
                  acc.x += val[j].x;
                  acc.y += val[j].y;
                  acc.z += val[j].z;
                  acc.w += val[j].w;
                }
              }
            }
          }
          if (valid) {
            float* out = (whole_rows || sseg == 0) ? A + sm.long_row[k] * A_LD : sm.extra + (size_t)(sm.long_x[k] + sseg - 1) * D;
            *reinterpret_cast<float4*>(out + 4 * l16) = acc;
          }
          sidx += 2 * NGW;
        }
        bar_sync(BAR_GATH, IT_GATH);
        if (tid == 0) sm.nlong = 0, sm.nxtra = 0;  // everybody has read them; the next tile registers after the last barrier
        if (!whole_rows) {
          for (int k2 = hw; k2 < nlong; k2 += 2 * NGW) {  // fold in segment order
            const int r = sm.long_row[k2];
            float4 a4 = *reinterpret_cast<const float4*>(A + r * A_LD + 4 * l16);
            const int nx = (sm.long_n[k2] - 1) / LONG, xb = sm.long_x[k2];
            for (int q = 0; q < nx; ++q) {
              const float4 x = *reinterpret_cast<const float4*>(sm.extra + (size_t)(xb + q) * D + 4 * l16);
              a4.x += x.x;
              a4.y += x.y;
              a4.z += x.z;
              a4.w += x.w;
            }
            *reinterpret_cast<float4*>(A + r * A_LD + 4 * l16) = a4;
          }
        }
        bar_sync(BAR_GATH, IT_GATH);
      }
      __threadfence_block();
      bar_arrive(BAR_FULL0 + buf, IT_T);  // hand the tile to the consumers
      ++it;
    }
    for (int j = max(0, it - 2); j < it; ++j) bar_sync(BAR_EMPTY0 + (j & 1), IT_T);  // drain
  } else {
    // ===================================== CONSUMER =====================================
    const int ct = tid - IT_GATH;
    const int tx = ct & 15, ty = ct >> 4;
    for (int i = ct; i < D * D / 4; i += IT_CONS) {
      reinterpret_cast<float4*>(sm.Wm)[i] = reinterpret_cast<const float4*>(W.w13)[i];
      reinterpret_cast<float4*>(sm.Wm + D * D)[i] = reinterpret_cast<const float4*>(W.w23)[i];
    }
    bar_sync(BAR_CONS, IT_CONS);
    int it = 0;
    for (; ti.next(); ti.tile += gridDim.x, ++it) {
      const int buf = it & 1;
      float* __restrict__ Hout = K.h[(iter + 1) & 1][ti.l];
      const float* A = sm.A[buf];
      bar_sync(BAR_FULL0 + buf, IT_T);
      // per-tile partial of the subgraph sum (old embeddings), fixed order => deterministic
      {
        const int c = ct & 63, part = ct >> 6;  // 4 row quarters x 64 columns
        float s = 0.f;
        for (int r = part * 32; r < part * 32 + 32; ++r) s += A[r * A_LD + D + c];
        sm.red[buf][part * D + c] = s;
      }
      bar_sync(BAR_CONS, IT_CONS);
      if (ct < D) {
        const float* red = sm.red[buf];
        const float s = ((red[ct] + red[D + ct]) + red[2 * D + ct]) + red[3 * D + ct];
        K.psum[(((size_t)iter * G.n_layers + ti.l) * ti.tiles_per_layer + ti.gt) * D + ct] = s;
      }
      if (dbg & 2) {
        __threadfence_block();
        bar_arrive(BAR_EMPTY0 + buf, IT_T);
        continue;
      }
      float acc[8][4];
      tile_gemm_128(A, sm.Wm, ty, tx, acc);
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

}  // namespace md

int g_md_debug_flags = 0;
extern "C" void md_debug_set_flags(int32_t f) { g_md_debug_flags = f; }

int md_embed_iter_tc_launch(const md_graphs_t* g, const md_state_t* s, const md_weights_t* w, const md_work_t* k, int iter,
                            cudaStream_t st, int* launched);

extern "C" int md_embed_iter(const md_graphs_t* g, const md_state_t* s, const md_weights_t* w, const md_work_t* k,
                             int32_t iter, md_stream_t stream) {
  if (!g || !s || !w || !k) return MD_EINVAL;
  if (g->n_layers < 1 || g->n_layers > MD_MAX_LAYERS) return MD_ELAYERS;
  if (g->stride % MD_TILE_M != 0 || g->n_total != g->n_graphs * g->stride) return MD_EINVAL;
  if (iter < 0 || iter >= MD_BP_ITERS) return MD_EINVAL;
  {  // tensor-core variant (md_embed_iter_tc.cu) only when md_debug_set_flags bit 4 asks for it (no faster: gather bound)
    int launched = 0;
    const int rc = md_embed_iter_tc_launch(g, s, w, k, iter, (cudaStream_t)stream, &launched);
    if (rc) return rc;
    if (launched) return 0;
  }
  static md::PerDevice<bool> configured_pd;
  bool& configured = configured_pd.get();
  static md::PerDevice<int> nsm_pd;
  int& nsm = nsm_pd.get();
  const int smem = (int)sizeof(md::IterSmem);
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(md::embed_iter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    if (nsm <= 0) nsm = 148;
    configured = true;
  }
  const int n_tiles = g->n_graphs * (g->stride / MD_TILE_M) * g->n_layers;
  int grid = nsm < n_tiles ? nsm : n_tiles;
  md::embed_iter_kernel<<<grid, md::IT_T, smem, (cudaStream_t)stream>>>(*g, *s, *w, *k, iter, g_md_debug_flags);
  MD_CHECK_LAUNCH();
  return 0;
}
