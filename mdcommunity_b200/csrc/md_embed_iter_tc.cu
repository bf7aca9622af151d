// md_embed_iter on the tensor cores: the unfused message-passing iteration (graphs too large for md_embed_fused, i.e. the
// N = 1M configuration) with the dense transform moved from FP32 FMA to tcgen05.
// Same reference lines as md_embed_iter.cu (MultiDismantler_net_graphsage.py:288-310) and the same warp-autonomous
// neighbour gather (descriptor prefetch one tile ahead, barrier-free short rows, cooperative hub rows); what changes:
//   * the 16 gather warps write the tile's operands straight into tcgen05 SWIZZLE_128B K-major tiles: P = A_adj.H and the
//     own rows H as raw fp32 (two 32-column K blocks each) plus their fp16 remainders (x - tf32(x));
//   * one MMA warp issues the 3-term product [P|H] @ [W13;W23] (tf32 hi x hi, hi x lo, fp16 remainder pass: 40 small MMAs
//     per tile) into one of two TMEM accumulator stages and releases the operand tiles as soon as they have been read;
//   * 4 epilogue warps (one thread per row) read the accumulators, relu, L2 normalise and store H' while the gather of
//     the next tile is already running.  With the FP32-FMA consumers gone the gather no longer shares issue slots with
//     a GEMM (measured on the FMA kernel: gather alone 298 us, GEMM alone 269 us, together 417 us per launch at C2).
#include <cuda_fp16.h>
#include <type_traits>
#include "md_common.cuh"
#include "md_umma.cuh"

namespace md {
using namespace umma;

constexpr int IT_GATH = 512;  // 16 gather warps
constexpr int NGW = IT_GATH / 32;
constexpr int TC_T = IT_GATH + 32 + 256;  // + MMA warp (16) + 8 epilogue warps (17..24): two per TMEM lane quadrant
constexpr int LONG = 16;      // rows with more alive neighbours are handled cooperatively
constexpr int WSTG = 8 * LONG;
constexpr int MAXLONG = TILE_M;
constexpr int MAXXS = 64;     // extra (2nd..) segments per tile; more -> whole-row fallback
constexpr int NACT_CACHE = 256;
constexpr uint32_t KB16 = 16384;

struct IterTcSmem {
  unsigned char P[2][KB16];    // neighbour sums, raw fp32, K blocks of 32 columns
  unsigned char H[2][KB16];    // own rows, raw fp32
  unsigned char LO[2][KB16];   // fp16 remainders: block 0 = P_lo (64 columns), block 1 = H_lo
  unsigned char Wb[5][KB16];   // B operands: W13 hi, W13 lo, W23 hi, W23 lo (2 K blocks of 8 KB each), fp16 [W13 | W23]
  float extra[MAXXS * D];      // partial sums of the 2nd.. segments of long rows
  float cs[2][NGW][D];         // [tile parity] per-warp column sums of the tile's own rows
  int wstg[2][NGW][WSTG];
  int wptr[2][NGW][12];
  int wv[2][NGW][8];
  int v[2][TILE_M];            // global node id of the row, -1 = padding row
  int nact[NACT_CACHE];
  int long_row[MAXLONG];
  int long_beg[MAXLONG];
  int long_n[MAXLONG];
  int long_x[MAXLONG];
  int nlong, nxtra;
  uint64_t a_full, a_empty, d_full[2], d_empty[2];
  uint32_t tmem;
  alignas(16) float h0_x[64], h0_inv[64];  // PRE, iteration 0 from the degrees: x = d / dmax and 1 / |relu([x, x] @ w_n2l)| per alive degree d
                               // PRE, iteration 1 from the degrees (H1 mode): a = c @ W13 and b = c @ W23
  int h0_key;                  // (layer, graph, dmax) the two tables were built for
  float c0[64];                // H1 mode: the common first embedding c = relu(w0 + w1) / |relu(w0 + w1)|
  int cnt[2][NGW];             // H1 mode: [tile parity] rows of the tile with an alive neighbour, per gather warp
};
static_assert(sizeof(IterTcSmem) <= 232448, "IterTcSmem exceeds the shared memory of one CTA");

enum { BAR_GATH = 5 };
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ bool it_elect() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xffffffff;\n\t@px mov.s32 %0, 1;\n\t}" : "+r"(pred));
  return pred != 0;
}

struct TileIterTc {  // the tile sequence all roles walk in lock step
  int tile, n_tiles, tiles_per_layer, tpg;
  int l, gt, g, t, nact;
  const int* nact_tab;
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

// PRE: the neighbour sums were computed by md_agg (agg_rows_kernel) into the rows of the output buffer; the gather warps only
// stream P and H rows in (contiguous 256-byte rows, all loads of a tile in flight before the operand tiles are touched).
template <bool PRE>
__global__ void __launch_bounds__(TC_T, 1) embed_iter_tc_kernel(md_graphs_t G, md_state_t S, md_weights_t W, md_work_t K, int iter, int dbg) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  IterTcSmem& sm = *reinterpret_cast<IterTcSmem*>(smem_raw);
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
  const bool cached = G.n_graphs <= NACT_CACHE;
  if (cached)
    for (int i = tid; i < G.n_graphs; i += TC_T) sm.nact[i] = S.n_act[i];
  if (tid == 0) {
    sm.nlong = 0, sm.nxtra = 0;
    sm.h0_key = -1;
    mbar_init(&sm.a_full, NGW);
    mbar_init(&sm.a_empty, 1);
    mbar_init(&sm.d_full[0], 1);
    mbar_init(&sm.d_full[1], 1);
    mbar_init(&sm.d_empty[0], 8);
    mbar_init(&sm.d_empty[1], 8);
    mbar_fence_init();
  }
  if (warp == NGW) tmem_alloc(&sm.tmem, 128);
  // weight operand images: element (k, n) of W13 / W23 -> row n of K block k / 32, tf32 hi and lo planes + fp16 plane
  {
    float t13[13], t23[13];
#pragma unroll
    for (int j = 0; j < 13; ++j) {
      const int i = tid + TC_T * j;
      t13[j] = i < D * D ? W.w13[i] : 0.f;
      t23[j] = i < D * D ? W.w23[i] : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 13; ++j) {
      const int i = tid + TC_T * j;
      if (i < D * D) {
        const int k = i >> 6, n = i & 63, kb = k >> 5, kk = k & 31;
        const uint32_t off = kb * 8192 + sw128_off(n, kk >> 2) + 4 * (kk & 3);
        *reinterpret_cast<float*>(sm.Wb[0] + off) = tf32_hi(t13[j]);
        *reinterpret_cast<float*>(sm.Wb[1] + off) = t13[j] - tf32_hi(t13[j]);
        *reinterpret_cast<float*>(sm.Wb[2] + off) = tf32_hi(t23[j]);
        *reinterpret_cast<float*>(sm.Wb[3] + off) = t23[j] - tf32_hi(t23[j]);
        *reinterpret_cast<__half*>(sm.Wb[4] + sw128_off(n, k >> 3) + 2 * (k & 7)) = __float2half_rn(t13[j]);
        *reinterpret_cast<__half*>(sm.Wb[4] + 8192 + sw128_off(n, k >> 3) + 2 * (k & 7)) = __float2half_rn(t23[j]);
      }
    }
  }
  if (PRE && (dbg & (1 << 21))) {
    // H1 mode (unit cost, iteration 1; see agg_deg_kernel<1> in md_agg.cu): the embedding after iteration 0 is a function of
    // the alive degree only, H1(d) = normalise(relu(d a + b)), a = c @ W13, b = c @ W23; table of d <= 63 in sm.extra
    // (unused in PRE mode), a / b kept for the hub rows beyond it
    if (warp == 0) {
      const float w0a = W.w_n2l[2 * lane], w0b = W.w_n2l[2 * lane + 1], w1a = W.w_n2l[D + 2 * lane], w1b = W.w_n2l[D + 2 * lane + 1];
      const float ca = fmaxf(fmaf(1.f, w1a, w0a), 0.f), cb = fmaxf(fmaf(1.f, w1b, w0b), 0.f);
      float ss = ca * ca + cb * cb;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      const float den = fmaxf(sqrtf(ss), 1e-12f);
      sm.c0[2 * lane] = ca / den, sm.c0[2 * lane + 1] = cb / den;
    }
    __syncthreads();
    if (tid < D) {
      float a = 0.f, b = 0.f;
      for (int k = 0; k < D; ++k) {
        a = fmaf(sm.c0[k], W.w13[k * D + tid], a);
        b = fmaf(sm.c0[k], W.w23[k * D + tid], b);
      }
      sm.h0_x[tid] = a, sm.h0_inv[tid] = b;
    }
    __syncthreads();
    for (int d = warp; d < 64; d += TC_T / 32) {
      const float fd = (float)d;
      const float x = d > 0 ? fmaxf(fmaf(fd, sm.h0_x[lane], sm.h0_inv[lane]), 0.f) : 0.f;
      const float y = d > 0 ? fmaxf(fmaf(fd, sm.h0_x[lane + 32], sm.h0_inv[lane + 32]), 0.f) : 0.f;
      float ss = x * x + y * y;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
      sm.extra[d * D + lane] = x * inv, sm.extra[d * D + lane + 32] = y * inv;
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = sm.tmem;
  TileIterTc ti;
  ti.tpg = G.stride / TILE_M;
  ti.tiles_per_layer = G.n_graphs * ti.tpg;
  ti.n_tiles = ti.tiles_per_layer * G.n_layers;
  ti.tile = blockIdx.x;
  ti.nact_tab = cached ? sm.nact : S.n_act;
  int n_tiles_done = 0;

  if (tid < IT_GATH) {
    // byte offset of the 16-byte chunk holding columns 4*l16.. of row r inside its 32-column K block
    auto p_off = [](int r, int c16) -> uint32_t { return (uint32_t)r * 128u + (uint32_t)(((c16 & 7) ^ (r & 7)) << 4); };
    const int l16_ = lane & 15;
    auto lo4 = [](const float4& x) -> uint2 {
      const __half2 a = __floats2half2_rn(x.x - tf32_hi(x.x), x.y - tf32_hi(x.y));
      const __half2 b = __floats2half2_rn(x.z - tf32_hi(x.z), x.w - tf32_hi(x.w));
      return make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
    };
    // this lane's float4 (columns 4*l16..) of row r: raw value + fp16 remainder (64 fp16 per row = one 128-byte row)
    auto st_p = [&](int r, const float4& x) {
      *reinterpret_cast<float4*>(sm.P[l16_ >> 3] + p_off(r, l16_)) = x;
      *reinterpret_cast<uint2*>(sm.LO[0] + (uint32_t)r * 128u + (((l16_ >> 1) ^ (r & 7)) << 4) + 8 * (l16_ & 1)) = lo4(x);
    };
    auto st_h = [&](int r, const float4& x) {
      *reinterpret_cast<float4*>(sm.H[l16_ >> 3] + p_off(r, l16_)) = x;
      *reinterpret_cast<uint2*>(sm.LO[1] + (uint32_t)r * 128u + (((l16_ >> 1) ^ (r & 7)) << 4) + 8 * (l16_ & 1)) = lo4(x);
    };
    // ===================================== GATHER WARPS =====================================
    const int half = lane >> 4, l16 = lane & 15;
    const int hw = warp * 2 + half;
    const int r0 = warp * 8;  // this warp's rows of every tile

    // descriptor prefetch of a tile into staging slot `par`: row pointers, node ids, neighbour ids of short rows
    auto prefetch = [&](const TileIterTc& tt, int par) {
      const int rows = min(TILE_M, tt.nact - tt.t * TILE_M);
      if constexpr (PRE) {
        const int v = (lane < 8 && r0 + lane < rows) ? S.act_list[tt.g * G.stride + tt.t * TILE_M + r0 + lane] : -1;
        if (lane < 8) sm.wv[par][warp][lane] = v;
        // start the DRAM -> L2 fetch of the next tile's P and H rows now (one 128-byte line per lane: 8 rows x 2 operands x
        // 2 lines); the demand loads follow one tile period later and find the lines in L2
        if (!(dbg & 2048)) {
          const int vv = __shfl_sync(0xffffffffu, v, lane >> 2);
          if (vv >= 0 && !((lane & 2) && (dbg & (3 << 20)))) {  // (no H rows to fetch when they come from the degrees)
            const float* base = (lane & 2) ? K.h[iter & 1][tt.l] : K.h[(iter + 1) & 1][tt.l];
            asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (size_t)vv * D + 32 * (lane & 1)));
          }
        }
        return;
      }
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
        if (lane < nr) sm.wstg[par][warp][dst0 + lane] = acol[src0 + lane];
      }
    };

    int it = 0;
    bool have = ti.next();
    if (have) prefetch(ti, 0);
    while (have) {
      const int buf = it & 1;
      const TileIterTc cur = ti;
      // advance to the next tile and prefetch its descriptors (other parity) before touching this tile's rows
      ti.tile += gridDim.x;
      have = ti.next();
      if constexpr (!PRE) {
        if (have) prefetch(ti, buf ^ 1);
        __syncwarp();
      }

      const int l = cur.l, g = cur.g, t = cur.t;
      const int rows = min(TILE_M, cur.nact - t * TILE_M);
      const int32_t* __restrict__ acol = S.acol[l];
      const float4* __restrict__ H4 = reinterpret_cast<const float4*>(K.h[iter & 1][l]);
      float4 csum = make_float4(0.f, 0.f, 0.f, 0.f);
      if constexpr (PRE) {
        const float4* __restrict__ P4 = reinterpret_cast<const float4*>(K.h[(iter + 1) & 1][l]);
        float4 pv[4], own[4];
        int vr[4];
        const bool h0_from_deg = (dbg & (1 << 20)) != 0;  // iteration 0, unit cost: H0 was never written (see md_agg.cu)
        const bool h1_from_deg = (dbg & (1 << 21)) != 0;  // iteration 1, unit cost: neither was H1
        int dg[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {  // all 8 row loads of this lane are in flight before anything is waited for
          vr[i] = sm.wv[buf][warp][2 * i + half];
          pv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          own[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          dg[i] = 0;
          if (vr[i] >= 0) {
            pv[i] = __ldcg(P4 + (size_t)vr[i] * (D / 4) + l16);  // other rows of this buffer are being overwritten: no .nc
            if (h0_from_deg || h1_from_deg) dg[i] = S.deg[l][vr[i]];
            else own[i] = __ldcg(H4 + (size_t)vr[i] * (D / 4) + l16);  // .cg: with 216 KB of shared memory the L1 is ~30 KB
          }
        }
        if (h0_from_deg) {  // relu([x, x] @ w_n2l) normalised, x = alive degree / max alive degree (embed_init_kernel's formula)
          // x and the reciprocal norm depend on the degree only: tables per (layer, graph), rebuilt by the 16 loader warps
          // with embed_init_kernel's warp-per-row arithmetic when the tile sequence moves on to another layer / graph
          // (computing them per row in every lane -- 4 divisions, square roots and reciprocals per lane and tile --
          // made this launch 0.2 ms slower than the ones that load H).
          const int dmax_i = S.max_deg[l * G.n_graphs + g];
          const int key = (l * G.n_graphs + g) * 1024 + min(dmax_i, 1023);
          if (sm.h0_key != key) {  // (uniform over the loader warps: they walk the same tiles)
            bar_sync(BAR_GATH, IT_GATH);
            const float w0a = W.w_n2l[2 * lane], w0b = W.w_n2l[2 * lane + 1];
            const float w1a = W.w_n2l[D + 2 * lane], w1b = W.w_n2l[D + 2 * lane + 1];
            for (int d = warp; d < 64; d += NGW) {
              const float x0 = (float)d / (float)dmax_i;
              const float a = fmaxf(fmaf(x0, w1a, x0 * w0a), 0.f), b = fmaxf(fmaf(x0, w1b, x0 * w0b), 0.f);
              float ss = a * a + b * b;
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
              if (lane == 0) sm.h0_x[d] = x0, sm.h0_inv[d] = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
            }
            if (tid == 0) sm.h0_key = key;
            bar_sync(BAR_GATH, IT_GATH);
          }
          const float dmax = (float)dmax_i;
          const float4 w0 = *reinterpret_cast<const float4*>(W.w_n2l + 4 * l16), w1 = *reinterpret_cast<const float4*>(W.w_n2l + D + 4 * l16);
          bool big = false;
#pragma unroll
          for (int i = 0; i < 4; ++i) big |= dg[i] > 63;
          big = __any_sync(0xffffffffu, big);  // hub rows beyond the tables: the whole warp takes the direct formula
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float x0, inv = 0.f;
            if (!big) x0 = sm.h0_x[dg[i]], inv = sm.h0_inv[dg[i]];
            else x0 = (float)dg[i] / dmax;
            float4 a;
            a.x = fmaxf(fmaf(x0, w1.x, x0 * w0.x), 0.f);
            a.y = fmaxf(fmaf(x0, w1.y, x0 * w0.y), 0.f);
            a.z = fmaxf(fmaf(x0, w1.z, x0 * w0.z), 0.f);
            a.w = fmaxf(fmaf(x0, w1.w, x0 * w0.w), 0.f);
            if (big) {
              float ss = (a.x * a.x + a.y * a.y) + (a.z * a.z + a.w * a.w);
              ss += __shfl_xor_sync(0xffffffffu, ss, 1);
              ss += __shfl_xor_sync(0xffffffffu, ss, 2);
              ss += __shfl_xor_sync(0xffffffffu, ss, 4);
              ss += __shfl_xor_sync(0xffffffffu, ss, 8);
              inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
            }
            own[i] = make_float4(a.x * inv, a.y * inv, a.z * inv, a.w * inv);
          }
        }
        if (h1_from_deg) {
          bool big = false;
          int c0n = 0;
#pragma unroll
          for (int i = 0; i < 4; ++i) big |= dg[i] > 63, c0n += (l16 == 0 && dg[i] >= 1) ? 1 : 0;
          big = __any_sync(0xffffffffu, big);  // hub rows beyond the table: the whole warp takes the direct formula
          const float4 a4 = *reinterpret_cast<const float4*>(&sm.h0_x[4 * l16]), b4 = *reinterpret_cast<const float4*>(&sm.h0_inv[4 * l16]);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (!big) {
              own[i] = *reinterpret_cast<const float4*>(&sm.extra[dg[i] * D + 4 * l16]);
            } else {
              const float fd = (float)dg[i];
              float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
              if (dg[i] > 0) z = make_float4(fmaxf(fmaf(fd, a4.x, b4.x), 0.f), fmaxf(fmaf(fd, a4.y, b4.y), 0.f),
                                             fmaxf(fmaf(fd, a4.z, b4.z), 0.f), fmaxf(fmaf(fd, a4.w, b4.w), 0.f));
              float ss = (z.x * z.x + z.y * z.y) + (z.z * z.z + z.w * z.w);
              ss += __shfl_xor_sync(0xffffffffu, ss, 1);
              ss += __shfl_xor_sync(0xffffffffu, ss, 2);
              ss += __shfl_xor_sync(0xffffffffu, ss, 4);
              ss += __shfl_xor_sync(0xffffffffu, ss, 8);
              const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
              own[i] = make_float4(z.x * inv, z.y * inv, z.z * inv, z.w * inv);
            }
          }
          // iteration 0's column sums (the graph row's input) were never formed either: sum_v H0[v] = (rows with an alive neighbour) c
          c0n = warp_sum_i(c0n);
          if (lane == 0) sm.cnt[buf][warp] = c0n;
        }
        if (have) prefetch(ti, buf ^ 1);  // node ids of the next tile: behind the row loads, off the critical path
        if (it >= 2) mbar_wait(&sm.d_empty[buf], ((it >> 1) - 1) & 1);  // epilogue of tile it-2 is done with v[buf]
        if (it >= 1) mbar_wait(&sm.a_empty, (it - 1) & 1);              // MMAs of the previous tile have read the operand tiles
        if (lane < 8) sm.v[buf][r0 + lane] = sm.wv[buf][warp][lane];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          st_p(r0 + 2 * i + half, pv[i]);
          st_h(r0 + 2 * i + half, own[i]);
          csum.x += own[i].x, csum.y += own[i].y, csum.z += own[i].z, csum.w += own[i].w;
        }
      } else {
      if (it >= 2) mbar_wait(&sm.d_empty[buf], ((it >> 1) - 1) & 1);  // epilogue of tile it-2 is done with v[buf]
      if (it >= 1) mbar_wait(&sm.a_empty, (it - 1) & 1);              // MMAs of the previous tile have read the operand tiles

      const int* wp = sm.wptr[buf][warp];
      const int* ws = sm.wstg[buf][warp];
      const int base_e = wp[9];
      if (lane < 8) sm.v[buf][r0 + lane] = sm.wv[buf][warp][lane];
      __syncwarp();
      // ---- short rows: 4 row pairs, one row per half-warp, no barrier --------------------------------------
      int spos = 0;  // running position in the packed staging area (same in both halves: recomputed per pair)
      {
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
          if (!is_long) st_p(r0 + 2 * i + half, acc);
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
        for (int i = 0; i < 4; ++i) {
          st_h(r0 + 2 * i + half, own[i]);
          csum.x += own[i].x, csum.y += own[i].y, csum.z += own[i].z, csum.w += own[i].w;
        }
      }
      // ---- long (hub) rows: cooperative, only when the tile has any ---------------------------------------------
      bar_sync(BAR_GATH, IT_GATH);
      const int nlong = sm.nlong;
      const bool whole_rows = sm.nxtra > MAXXS;  // too many segments for the scratch area: one half-warp per long row
      if (nlong > 0) {
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
            if (whole_rows) st_p(sm.long_row[k], acc);
            else if (sseg == 0) *reinterpret_cast<float4*>(sm.P[l16 >> 3] + p_off(sm.long_row[k], l16)) = acc;  // raw; LO after the fold
            else *reinterpret_cast<float4*>(sm.extra + (size_t)(sm.long_x[k] + sseg - 1) * D + 4 * l16) = acc;
          }
          sidx += 2 * NGW;
        }
        bar_sync(BAR_GATH, IT_GATH);
        if (tid == 0) sm.nlong = 0, sm.nxtra = 0;  // everybody has read them; the next tile registers after the last barrier
        if (!whole_rows) {
          for (int k2 = hw; k2 < nlong; k2 += 2 * NGW) {  // fold in segment order
            const int r = sm.long_row[k2];
            float4 a4 = *reinterpret_cast<const float4*>(sm.P[l16 >> 3] + p_off(r, l16));
            const int nx = (sm.long_n[k2] - 1) / LONG, xb = sm.long_x[k2];
            for (int q = 0; q < nx; ++q) {
              const float4 x = *reinterpret_cast<const float4*>(sm.extra + (size_t)(xb + q) * D + 4 * l16);
              a4.x += x.x;
              a4.y += x.y;
              a4.z += x.z;
              a4.w += x.w;
            }
            st_p(r, a4);
          }
        }
        bar_sync(BAR_GATH, IT_GATH);
      }
      }  // !PRE
      // per-tile column sums of the old embeddings (spmm subgsum): half-warps, then warps, in a fixed order
      csum.x += __shfl_xor_sync(0xffffffffu, csum.x, 16);
      csum.y += __shfl_xor_sync(0xffffffffu, csum.y, 16);
      csum.z += __shfl_xor_sync(0xffffffffu, csum.z, 16);
      csum.w += __shfl_xor_sync(0xffffffffu, csum.w, 16);
      if (lane < 16) *reinterpret_cast<float4*>(&sm.cs[buf][warp][4 * l16]) = csum;
      bar_sync(BAR_GATH, IT_GATH);
      if (tid < D) {
        float s = 0.f;
#pragma unroll
        for (int w2 = 0; w2 < NGW; ++w2) s += sm.cs[buf][w2][tid];
        K.psum[(((size_t)iter * G.n_layers + l) * cur.tiles_per_layer + cur.gt) * D + tid] = s;
        if (PRE && (dbg & (1 << 21))) {
          int n0 = 0;
#pragma unroll
          for (int w2 = 0; w2 < NGW; ++w2) n0 += sm.cnt[buf][w2];
          K.psum[(((size_t)0 * G.n_layers + l) * cur.tiles_per_layer + cur.gt) * D + tid] = (float)n0 * sm.c0[tid];
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.a_full);  // hand the operand tiles to the MMA warp
      ++it;
    }
    n_tiles_done = it;
  } else if (warp == NGW) {
    // ===================================== MMA ISSUER =====================================
    const uint32_t idt = idesc(FMT_TF32, 128, 64), idf = idesc(FMT_F16, 128, 64);
    uint64_t aP[2], aH[2], aLO[2], b13h[2], b13l[2], b23h[2], b23l[2], bf[2];
#pragma unroll
    for (int kb = 0; kb < 2; ++kb) {
      aP[kb] = desc_k_sw128(smem_u32(sm.P[kb]), 0);
      aH[kb] = desc_k_sw128(smem_u32(sm.H[kb]), 0);
      aLO[kb] = desc_k_sw128(smem_u32(sm.LO[kb]), 0);
      b13h[kb] = desc_k_sw128(smem_u32(sm.Wb[0]) + kb * 8192, 0);
      b13l[kb] = desc_k_sw128(smem_u32(sm.Wb[1]) + kb * 8192, 0);
      b23h[kb] = desc_k_sw128(smem_u32(sm.Wb[2]) + kb * 8192, 0);
      b23l[kb] = desc_k_sw128(smem_u32(sm.Wb[3]) + kb * 8192, 0);
      bf[kb] = desc_k_sw128(smem_u32(sm.Wb[4]) + kb * 8192, 0);
    }
    uint32_t it = 0;
    for (; ti.next(); ti.tile += gridDim.x, ++it) {
      const uint32_t stage = it & 1;
      mbar_wait(&sm.a_full, it & 1);
      if (it >= 2) mbar_wait(&sm.d_empty[stage], ((it >> 1) - 1) & 1);  // accumulator stage drained by the epilogue
      tc_fence_after();
      if (it_elect()) {
        const uint32_t d = tbase + 64 * stage;
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t ko = 2 * k;
            mma_tf32(d, aP[kb] + ko, b13h[kb] + ko, idt, (kb | k) != 0);
            mma_tf32(d, aP[kb] + ko, b13l[kb] + ko, idt, 1);
            mma_tf32(d, aH[kb] + ko, b23h[kb] + ko, idt, 1);
            mma_tf32(d, aH[kb] + ko, b23l[kb] + ko, idt, 1);
            mma_f16(d, aLO[kb] + ko, bf[kb] + ko, idf, 1);
          }
        mma_commit(&sm.a_empty);
        mma_commit(&sm.d_full[stage]);
      }
      __syncwarp();
    }
  } else {
    // ===================================== EPILOGUE: one thread per row and column half =====================
    // Two warps share a TMEM lane quadrant (warp % 4): both read all 64 accumulators of their row for the norm, each
    // writes 32 of the 64 outputs.  (With one warp per quadrant the 16 scattered 16-byte stores per thread made the
    // epilogue the longest stage of the PRE pipeline.)
    const int quad = warp & 3, chalf = (warp - NGW - 1) >> 2;
    const int row = 32 * quad + lane;
    auto epilogue = [&](auto ch_tag) {
      constexpr int CH = decltype(ch_tag)::value;  // column half, compile-time so that `keep` stays in registers
      uint32_t it = 0;
      for (; ti.next(); ti.tile += gridDim.x, ++it) {
        const uint32_t stage = it & 1;
        float* __restrict__ Hout = K.h[(iter + 1) & 1][ti.l];
        mbar_wait(&sm.d_full[stage], (it >> 1) & 1);
        tc_fence_after();
        const uint32_t taddr = tbase + ((uint32_t)(32 * quad) << 16) + 64 * stage;
        const int v = sm.v[stage][row];
        float ss = 0.f;
        float keep[2][16];  // this warp's 32 columns, kept from the norm pass
#pragma unroll
        for (int c0 = 0; c0 < D; c0 += 16) {
          float x[16];
          tmem_ld16(taddr + c0, x);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float r = fmaxf(x[j], 0.f);
            ss = fmaf(r, r, ss);
            if ((c0 >> 5) == CH) keep[(c0 >> 4) & 1][j] = r;
          }
        }
        const float inv = __frcp_rn(fmaxf(sqrtf(ss), 1e-12f));
        if (v >= 0 && !(dbg & 1024)) {
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              *reinterpret_cast<float4*>(Hout + (size_t)v * D + 32 * CH + 16 * h + j) =
                  make_float4(keep[h][j] * inv, keep[h][j + 1] * inv, keep[h][j + 2] * inv, keep[h][j + 3] * inv);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.d_empty[stage]);
      }
    };
    if (chalf == 0) epilogue(std::integral_constant<int, 0>{});
    else epilogue(std::integral_constant<int, 1>{});
  }
  (void)n_tiles_done;
  tc_fence_before();
  __syncthreads();
  if (warp == NGW) tmem_dealloc(tbase, 128);
}

}  // namespace md

extern int g_md_debug_flags;

// *launched = 1 when a tensor-core kernel was enqueued.
//  * two-kernel path (default when the embeddings exceed L2, i.e. the N = 1M configuration; md_debug_set_flags bit 8 forces
//    it, bit 7 disables it): md_agg's high-occupancy neighbour aggregation, then this kernel in PRE mode.  C4: 9.0 -> see
//    DESIGN 6 ms per step for the aggregation stage.
//  * single fused tensor-core kernel: only on request (bit 4).  Measured: no gain over the FP32-FMA kernel (C2 unfused 450
//    vs 417 us per launch, C4 9.2 vs 9.0 ms per step): with the gather inside, the iteration is bound by the 16 gather
//    warps' row loads, which the FMA consumers overlapped for free; kept as the A/B reference of that finding.
int md_agg_rows_launch(const md_graphs_t* g, const md_state_t* s, const md_work_t* k, int iter, cudaStream_t st);
int md_agg_deg_launch(const md_graphs_t* g, const md_state_t* s, const md_weights_t* w, const md_work_t* k, int level, cudaStream_t st);

static bool tc_pre_path(const md_graphs_t* g) {
  const int f = g_md_debug_flags;
  if (f & (1 | 2)) return false;  // isolation runs of the FMA kernel
  const bool beyond_l2 = (size_t)g->n_total * g->n_layers * (md::D * 4) > ((size_t)96 << 20);
  return !(f & 128) && ((f & 256) || beyond_l2);
}
// true when the first embedding is never materialised: two-kernel path, unit cost (features = alive degree / max alive
// degree), not switched off by md_debug_set_flags bit 12.  md_embed_init asks this to become a no-op.
bool md_h0_on_the_fly(const md_graphs_t* g) { return tc_pre_path(g) && g->node_w[0] == nullptr && !(g_md_debug_flags & 4096); }
// true when the embedding after iteration 0 is not materialised either (md_agg.cu, agg_deg_kernel<1>): iteration 0 launches
// nothing, iteration 1 works from the alive degrees.  md_debug_set_flags bit 18 switches it off (iteration 0 then runs as
// agg_deg_kernel<0> + transform and leaves H1 in buffer 1: the A/B reference and what tests that look at H1 use).
static bool h1_on_the_fly(const md_graphs_t* g) { return md_h0_on_the_fly(g) && !(g_md_debug_flags & (1 << 18)); }

int md_embed_iter_tc_launch(const md_graphs_t* g, const md_state_t* s, const md_weights_t* w, const md_work_t* k, int iter,
                            cudaStream_t st, int* launched) {
  *launched = 0;
  const int f = g_md_debug_flags;
  if (f & (1 | 2)) return 0;  // isolation runs of the FMA kernel
  const bool pre = tc_pre_path(g);
  if (!pre && !(f & 16)) return 0;
  static md::PerDevice<bool> configured_pd;
  bool& configured = configured_pd.get();
  static md::PerDevice<int> nsm_pd;
  int& nsm = nsm_pd.get();
  const int smem = (int)sizeof(md::IterTcSmem);
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(md::embed_iter_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(md::embed_iter_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    if (nsm <= 0) nsm = 148;
    configured = true;
  }
  const int n_tiles = g->n_graphs * (g->stride / MD_TILE_M) * g->n_layers;
  const int grid = nsm < n_tiles ? nsm : n_tiles;
  if (pre) {
    const bool h1 = h1_on_the_fly(g);
    if (h1 && iter == 0) {  // nothing to do: H1 is a function of the alive degrees
      *launched = 1;
      return 0;
    }
    const bool h0 = iter == 0 && md_h0_on_the_fly(g);
    const bool h1it = h1 && iter == 1;
    // bits 14 / 15: enqueue only the aggregation / only the transform (bench.py times the two kernels of an iteration
    // separately by calling md_embed_iter twice; together the two calls are exactly the normal launch sequence)
    if (!(f & 32768)) {
      const int rc = (h0 || h1it) ? md_agg_deg_launch(g, s, w, k, h1it ? 1 : 0, st) : md_agg_rows_launch(g, s, k, iter, st);
      if (rc) return rc;
    }
    if (!(f & 16384))
      md::embed_iter_tc_kernel<true><<<grid, md::TC_T, smem, st>>>(*g, *s, *w, *k, iter, (f & 0x3fff) | (h0 ? (1 << 20) : 0) | (h1it ? (1 << 21) : 0));
  } else {
    md::embed_iter_tc_kernel<false><<<grid, md::TC_T, smem, st>>>(*g, *s, *w, *k, iter, f);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  *launched = 1;
  return 0;
}
