// md_embed_fused: the node-embedding stage of one forward (first embedding + the three message-passing iterations)
// for batches of small graphs (stride <= 1024), one launch, embeddings resident in shared memory.
// Replaces /root/reference/code/MultiDismantler_unit_cost/MultiDismantler_net_graphsage.py:248-310 (node features,
// relu(x @ w_n2l), normalise; 3 x { torch_sparse.spmm(n2nsum), the three dense transforms, relu, normalise,
// torch_sparse.spmm(subgsum) }) -- the same arithmetic md_embed_init + 3 x md_embed_iter perform through HBM/L2.
//
// Why a second formulation: ncu on md_embed_iter showed the neighbour gather pinned at the L2 -> SM path (one 256-byte
// row per alive edge, ~37 B/cycle/SM at best) with the FP32-FMA transform fighting it for issue slots.  For graphs of
// <= 1024 active nodes one graph-layer's embeddings are 256 KB: too large for one SM's shared memory in full, but one
// COLUMN HALF (32 of 64 columns, 128 B per row) is 128 KB.  So a persistent CTA owns one BIN at a time -- one (graph, layer)
// unit, or several small units of one layer packed together (fused_prep_kernel below):
//
//   per iteration, two passes p = 0,1 over the column halves:
//     resident  Hs = H[:, 32p:32p+32] in shared memory, compact row order, in the SWIZZLE_128B tile geometry (a 128-byte
//               row is one swizzle span: 8 lanes x LDS.128 per neighbour row, conflict-free).
//     gather    all 16 warps; 8 lanes per row (4 rows per warp instruction); neighbour lists come from a shared-memory copy of
//               the bin's alive CSR (uint16 codes, translated once per bin).  Rows are re-ordered per bin by descending
//               alive degree (deterministic counting sort) and dealt to the lane groups so that the rows a warp walks
//               together have (nearly) equal length.  Sums stay in CSR order; rows longer than FZ_LONG are spread over the
//               warp's four lane groups (first tile of a pass: taken from a counter as warps become free).
//     staging   per 128-row tile two tiles of fp16 pairs: [H_hi | H_lo] (the rows' own embeddings, staged first) and
//               [P_hi | P_lo] (P = A_adj.H), x = hi + lo, hi = fp16(x), lo = fp16(x - hi).
//     transform D_tile += P.W13[32p:32p+32] + H.W23[32p:32p+32] as a.w ~= a_hi.w_hi + a_hi.w_lo + a_lo.w_hi, all kind::f16 with
//               fp32 accumulation (12 MMAs of K = 16 per tile and pass; the weight images [W_hi | W_lo] of both passes stay
//               resident).  No MMA warp: the warps take turns waiting for a staging tile to be complete and issuing its MMAs
//               (a 17th warp would cost every thread registers: see FZ_T).  Accumulators for ALL tiles of the bin
//               (<= 8 x 64 columns = the whole 512-column TMEM) stay in tensor memory across the two passes, which is what
//               makes the column split free: no partial sums are exchanged.
//   epilogue    warps 0..7, tcgen05.ld (one thread per row holds the 64 outputs): relu, L2 normalise; columns 0..31 go
//               straight into the resident buffer for the next iteration's pass 0, columns 32..63 into a per-CTA scratch image
//               (L2-resident; one 4 KB bulk copy per warp and tile) that one cp.async.bulk brings back for pass 1.  The last
//               iteration writes H to HBM for md_qhead.  Per-graph column sums of every iteration's input (the subgsum spmm)
//               fall out of the gather (single units) or are formed member by member from the resident half (bins).
//
// HBM traffic per unit and forward: alive CSR once + final H once (+ features); everything else stays on chip / in L2.
#include <cuda_fp16.h>
#include "md_common.cuh"
#include "md_umma.cuh"

namespace md {
using namespace umma;

constexpr int FZ_ROWS = 1024;          // largest stride (node slots per graph) the resident path handles
#ifndef FZ_WARPS
#define FZ_WARPS 16
#endif
constexpr int FZ_GW = FZ_WARPS;        // warps of the CTA (8 or 16); all of them gather, warps 0..7 run the epilogue
constexpr int FZ_RPT = 1024 / (FZ_GW * 32);   // rows of a unit per thread (sort, first-embedding fill)
constexpr int FZ_RPG = 128 / (4 * FZ_GW);     // rows of a tile per lane group (gather)
static_assert(FZ_GW == 8 || FZ_GW == 16, "8 or 16 warps");
constexpr int FZ_GT = FZ_GW * 32;
constexpr int FZ_T = FZ_GT;            // no separate MMA warp: a ninth warp would cap every thread at 168 registers (three warps on
                                       // one of the SM's four register files), and at 168 ptxas serialises the gather's four load chains
constexpr int FZ_IDX_CAP = 8192;       // alive directed edges of one unit cached in shared memory
constexpr int FZ_LONG = 24;            // rows with more alive neighbours are summed by the whole warp (< 31: exact sort buckets)
constexpr uint32_t FZ_TILE_B = 16384;  // one 128-row x 128-byte tile
constexpr uint32_t FZ_WPASS_B = 16384; // B operands of one pass: [W13 hi | W13 lo] and [W23 hi | W23 lo], fp16, 64 rows x 128 B each
constexpr size_t FZ_WS_CTR = 2 * FZ_WPASS_B;
constexpr size_t FZ_WS_SPILL = FZ_WS_CTR + 1024;
constexpr size_t FZ_SPILL_B = 8 * FZ_TILE_B;
// ---- packing of small units (see fused_prep_kernel) ----
constexpr int FZ_MAXM = 32;          // most (graph, layer) units one CTA takes together (a "bin")
constexpr int FZ_SMALL = 256;        // a unit is packable when it has at most this many active rows ...
constexpr int FZ_SMALL_E = 2048;     // ... and at most this many alive entries (8 per row of FZ_SMALL)
constexpr int FZ_PACK_GRAPHS = 1024; // batches beyond this many graphs (or FZ_PACK_UNITS units) are not packed
constexpr int FZ_PACK_UNITS = 2048;
constexpr int FZ_CAP_MIN = 256, FZ_CAP_MAX = 768;  // bounds of a bin's weight budget (rows < budget + FZ_SMALL <= FZ_ROWS)

struct FusedBins {     // written by fused_prep_kernel, read by embed_fused_kernel
  int* hdr;            // [0] bin counter of the persistent CTAs, [1] number of bins
  int* order;          // [n_units] bins in the order the CTAs take them (heavy first)
  int* bin_start;      // [n_units + 1] bin b = member[bin_start[b] .. bin_start[b + 1])
  int* member;         // [n_units] unit ids (g * L + l); all members of a bin share the layer
  uint16_t* wpos;      // [L][n_total] packed bins: node slot -> J code of its row inside the bin (see nbrJ)
};

struct FusedSmem {
  unsigned char Hs[8 * FZ_TILE_B];  // resident column half of the unit's embeddings (SWIZZLE_128B tiles)
  unsigned char P[FZ_TILE_B];       // staging: neighbour sums of one tile as fp16 pairs, row = [P_hi (32) | P_lo (32)]
  unsigned char LO[FZ_TILE_B];      // staging: the rows' own embeddings the same way, row = [H_hi (32) | H_lo (32)]
  unsigned char Wt[2 * FZ_WPASS_B]; // B operands of both passes (loaded once per CTA)
  uint16_t idx[FZ_IDX_CAP + 64];    // neighbour entries of the unit's alive CSR (see nbrJ); setup: sort counters
  uint16_t pos[FZ_ROWS];            // single big unit: graph-local node id -> row (degree-sorted compact order);
                                    // packed bin: compact index (member row offset + p) -> J code of the row
  uint32_t rbe[FZ_ROWS];            // row -> first | (one-past-last << 16) entry of its neighbour list
  uint16_t rowv[FZ_ROWS];           // row -> graph-local node id | member index << 10
  uint16_t mg[FZ_MAXM];             // member -> graph
  uint16_t mrow[FZ_MAXM + 1];       // member -> first compact index (prefix sum of the members' active rows)
  uint32_t meo[FZ_MAXM + 1];        // member -> first entry inside the bin's concatenated neighbour lists
  int32_t me0[FZ_MAXM];             // member -> first entry of its alive CSR in S.acol
  int nm, nrows, nent, layer, packed;
  alignas(16) float red[FZ_GW][32];
  alignas(16) float wn2l[2 * D];     // first-embedding coefficients: relu(x1 * wn2l[64 + c] + x0 * wn2l[c]) (see H1 mode below)
  float c0[D];                      // H1 mode: the common first embedding c
  int cnt0;                         // H1 mode: rows of the unit with an alive neighbour
  alignas(128) float zrow[32];      // one all-zero 128-byte row: where a finished row's lane group points its loads
  uint16_t bbase[32];
  // per staging tile: "full" = every warp has staged its rows (FZ_GW arrivals), "empty" = the MMAs that read it have completed
  uint64_t bar_fullH, bar_fullP, bar_emptyH, bar_emptyP, bar_w, bar_h;
  int long_ctr;                     // tickets for the long rows of a pass's first tile
  uint32_t tmem;
  int unit;
};
static_assert(sizeof(FusedSmem) <= 232448, "FusedSmem exceeds the 227 KB shared memory of one CTA");

// phase clocks of CTA 0 (profiling): gather warp 0 -> [0..7], MMA warp -> [8..11], units processed -> [15]
__device__ long long g_fz_dbg[16];
__device__ unsigned g_fz_gath[8];  // gather sub-phases of CTA 0 warp 0: setup, main loop, own rows, stores, long rows, fence+arrive
__device__ unsigned g_fz_cta[160][2];  // per CTA: cycles from start to end of the unit loop, units processed
#ifdef MD_PROFILE
#define FZ_CLK(acc)                   \
  do {                                \
    const uint32_t n_ = clock();      \
    acc += n_ - t_dbg;                \
    t_dbg = n_;                       \
  } while (0)
#else
#define FZ_CLK(acc) do { } while (0)
#endif

__device__ __forceinline__ void gw_sync() { asm volatile("bar.sync 1, %0;" ::"n"(FZ_GT) : "memory"); }

__device__ __forceinline__ void add4(float4& a, const float4& b) {
  a.x += b.x;
  a.y += b.y;
  a.z += b.z;
  a.w += b.w;
}
// x = hi + lo with hi = fp16(x), lo = fp16(x - hi): 22 significant bits (|x| >= 2^-14; below that 2^-25 absolute)
__device__ __forceinline__ void split_f16x4(const float4& v, uint2& hi, uint2& lo) {
  const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
  const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
  const __half2 l0 = __floats2half2_rn(v.x - f0.x, v.y - f0.y), l1 = __floats2half2_rn(v.z - f1.x, v.w - f1.y);
  hi = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
  lo = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
}
// staging row `rt` of a [hi (32) | lo (32)] fp16 tile: lane q of the row's group owns columns 4q .. 4q + 3
__device__ __forceinline__ void stage_f16(unsigned char* tile, int rt, int q, const float4& v) {
  uint2 hi, lo;
  split_f16x4(v, hi, lo);
  const uint32_t ro = (uint32_t)rt * 128u + 8 * (q & 1);
  *reinterpret_cast<uint2*>(tile + ro + (((q >> 1) ^ (rt & 7)) << 4)) = hi;
  *reinterpret_cast<uint2*>(tile + ro + (((4 + (q >> 1)) ^ (rt & 7)) << 4)) = lo;
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xffffffff;\n\t@px mov.s32 %0, 1;\n\t}"
      : "+r"(pred));
  return pred != 0;
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
struct F4P {  // a float4 as two packed f32x2 registers (for FADD2)
  uint64_t lo, hi;
};
__device__ __forceinline__ F4P lds128p(uint32_t addr) {
  F4P v;
  asm volatile("ld.shared.v2.b64 {%0,%1}, [%2];" : "=l"(v.lo), "=l"(v.hi) : "r"(addr));
  return v;
}
__device__ __forceinline__ void add4p(F4P& a, const F4P& b) {
  asm("add.rn.f32x2 %0, %0, %1;" : "+l"(a.lo) : "l"(b.lo));
  asm("add.rn.f32x2 %0, %0, %1;" : "+l"(a.hi) : "l"(b.hi));
}
__device__ __forceinline__ float4 unpack4(const F4P& a) {
  return make_float4(__uint_as_float((uint32_t)a.lo), __uint_as_float((uint32_t)(a.lo >> 32)), __uint_as_float((uint32_t)a.hi),
                     __uint_as_float((uint32_t)(a.hi >> 32)));
}
__device__ __forceinline__ uint32_t lds16(uint32_t addr) {
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
  return v;
}

// Per launch, two blocks.  Block 1: the weight images (both passes) in the exact shared-memory layout.  Block 0: the BINS.
//
// Towards the end of a rollout most graphs of a batch have a few dozen active nodes left, and a (graph, layer) unit of 30
// rows costs a persistent CTA nearly as much as one of 300 (sort, fill, six passes with their barriers and MMA round trips,
// three epilogues: ~38 k cycles, measured).  So units of <= FZ_SMALL rows and <= FZ_SMALL_E alive entries are PACKED: a bin
// is a run of consecutive (in graph order) small units of one layer whose rows are concatenated into one resident buffer
// and go through the kernel like a single unit.  Nothing in a unit's arithmetic depends on its bin mates (neighbour sums
// run in CSR order, the tensor core and the epilogue work row by row, column sums run per member in the member's own
// order), so a graph's embeddings are bit-identical packed, alone, or in another batch.
//
// Packing is a prefix sum, not a search: unit weight w = max(rows, entries / 8, budget / FZ_MAXM); a unit whose exclusive
// prefix W falls into [k * budget, (k + 1) * budget) belongs to bin k of its layer.  Hence rows < budget + FZ_SMALL <=
// FZ_ROWS, entries < 8 (budget + FZ_SMALL) <= FZ_IDX_CAP, members <= FZ_MAXM.  budget = the small units' total weight
// spread over the SMs, rounded up to whole tiles, within [FZ_CAP_MIN, FZ_CAP_MAX] (md_debug_set_fused_pack overrides it;
// < 0: every unit alone).  Larger units get a bin each.  Bins are handed out heavy first (64 weight classes, counting
// sort; the order inside a class is whatever the atomics give -- it decides which CTA takes a bin, not any result).
__global__ void __launch_bounds__(1024) fused_prep_kernel(md_graphs_t G, md_state_t S, md_weights_t W, unsigned char* ws,
                                                          FusedBins bn, int n_units, int nsm, int pack) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (blockIdx.x == 1) {
    for (int i = tid; i < 2 * 64 * 32; i += 1024) {
      const int p = i >> 11, n = (i >> 5) & 63, k = i & 31;  // pass, output column (B operand row), input column of the half
      const float w13 = W.w13[(32 * p + k) * D + n], w23 = W.w23[(32 * p + k) * D + n];
      unsigned char* img = ws + (size_t)p * FZ_WPASS_B;
      const __half h13 = __float2half_rn(w13), h23 = __float2half_rn(w23);
      const uint32_t ohi = sw128_off(n, k >> 3) + 2 * (k & 7), olo = sw128_off(n, 4 + (k >> 3)) + 2 * (k & 7);
      *reinterpret_cast<__half*>(img + ohi) = h13;
      *reinterpret_cast<__half*>(img + olo) = __float2half_rn(w13 - __half2float(h13));
      *reinterpret_cast<__half*>(img + 8192 + ohi) = h23;
      *reinterpret_cast<__half*>(img + 8192 + olo) = __float2half_rn(w23 - __half2float(h23));
    }
    return;
  }
  const int L = G.n_layers, B = G.n_graphs, St = G.stride;
  if (B > FZ_PACK_GRAPHS || n_units > FZ_PACK_UNITS) {  // very large batches: one bin per unit, index order
    for (int u = tid; u < n_units; u += 1024) bn.bin_start[u] = u, bn.member[u] = u, bn.order[u] = u;
    if (tid == 0) bn.bin_start[n_units] = n_units, bn.hdr[0] = 0, bn.hdr[1] = n_units;
    return;
  }
  __shared__ int key[FZ_PACK_UNITS];                      // weight of a bin (hand-out order)
  __shared__ int lbst[MD_MAX_LAYERS][FZ_PACK_GRAPHS + 1];  // layer, local bin -> rank of its first member among the layer's small units
  __shared__ int wsum[3][32];
  __shared__ int tot[MD_MAX_LAYERS][3];                    // per layer: total weight, small units, big units
  __shared__ int snb[MD_MAX_LAYERS];                       // per layer: bins of small units
  __shared__ int hist[64], cursor[64];
  __shared__ int t_small;
  for (int i = tid; i < FZ_PACK_UNITS; i += 1024) key[i] = 0;
  for (int i = tid; i < MD_MAX_LAYERS * (FZ_PACK_GRAPHS + 1); i += 1024) (&lbst[0][0])[i] = 0x7fffffff;
  if (tid < 64) hist[tid] = 0;
  if (tid < MD_MAX_LAYERS) snb[tid] = 0;
  if (tid == 0) t_small = 0;
  __syncthreads();
  // ---- the thread's graph: rows, entries and class of its unit in every layer ----
  const int n = tid < B ? S.n_act[tid] : 0;
  int e_[MD_MAX_LAYERS], cls_[MD_MAX_LAYERS], w_[MD_MAX_LAYERS];  // class: 0 = nothing to do, 1 = small (packable), 2 = big
  int tsm = 0;
#pragma unroll
  for (int l = 0; l < MD_MAX_LAYERS; ++l) {
    e_[l] = 0, cls_[l] = 0, w_[l] = 0;
    if (l < L && n > 0) {
      e_[l] = S.aptr[l][tid * St + n] - S.aptr[l][tid * St];
      cls_[l] = (n <= FZ_SMALL && e_[l] <= FZ_SMALL_E) ? 1 : 2;
      w_[l] = max(n, (e_[l] + 7) >> 3);
      if (cls_[l] == 1) tsm += w_[l];
    }
  }
  tsm = warp_sum_i(tsm);
  if (lane == 0 && tsm) atomicAdd(&t_small, tsm);
  __syncthreads();
  int cap = 0;
  if (pack > 0) cap = min(max(pack, FZ_CAP_MIN), FZ_CAP_MAX);
  else if (pack == 0) cap = min(max(((t_small + nsm - 1) / nsm + TILE_M - 1) / TILE_M * TILE_M, FZ_CAP_MIN), FZ_CAP_MAX);
  const int wmin = (cap + FZ_MAXM - 1) / FZ_MAXM;
  // ---- exclusive prefixes over the graphs, per layer: weight, small units, big units ----
  int wex_[MD_MAX_LAYERS], srk_[MD_MAX_LAYERS], brk_[MD_MAX_LAYERS];
#pragma unroll
  for (int l = 0; l < MD_MAX_LAYERS; ++l) {
    wex_[l] = srk_[l] = brk_[l] = 0;
    if (l < L) {  // (L is uniform: the barriers below are reached by every thread)
      int v[3] = {cls_[l] == 1 ? max(w_[l], wmin) : 0, cls_[l] == 1 ? 1 : 0, cls_[l] == 2 ? 1 : 0};
      int inc[3];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        int x = v[j];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int y = __shfl_up_sync(0xffffffffu, x, o);
          if (lane >= o) x += y;
        }
        inc[j] = x;
        if (lane == 31) wsum[j][warp] = x;
      }
      __syncthreads();
      if (warp == 0) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int t = wsum[j][lane];
          int x = t;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
          }
          wsum[j][lane] = x - t;
          if (lane == 31) tot[l][j] = x;
        }
      }
      __syncthreads();
      wex_[l] = inc[0] - v[0] + wsum[0][warp];
      srk_[l] = inc[1] - v[1] + wsum[1][warp];
      brk_[l] = inc[2] - v[2] + wsum[2][warp];
      __syncthreads();  // wsum is reused by the next layer
    }
  }
  // ---- local bin of every small unit; first member and number of bins per layer ----
  int k_[MD_MAX_LAYERS];
#pragma unroll
  for (int l = 0; l < MD_MAX_LAYERS; ++l) {
    k_[l] = 0;
    if (l < L && cls_[l] == 1) {
      k_[l] = cap ? wex_[l] / cap : srk_[l];
      atomicMin(&lbst[l][k_[l]], srk_[l]);
      atomicMax(&snb[l], k_[l] + 1);
    }
  }
  __syncthreads();
  int n_big = 0, n_bins = 0, n_mem = 0;
#pragma unroll
  for (int l = 0; l < MD_MAX_LAYERS; ++l)
    if (l < L) n_big += tot[l][2];
  n_bins = n_mem = n_big;
  {
    int bigoff = 0;
#pragma unroll
    for (int l = 0; l < MD_MAX_LAYERS; ++l) {
      if (l < L) {
        const int boff = n_bins, moff = n_mem;  // first bin / first member slot of this layer's small units
        if (cls_[l] == 2) {
          const int b = bigoff + brk_[l];
          bn.member[b] = tid * L + l;
          bn.bin_start[b] = b;
          key[b] = e_[l] + 4 * n;
        } else if (cls_[l] == 1) {
          bn.member[moff + srk_[l]] = tid * L + l;
          atomicAdd(&key[boff + k_[l]], e_[l] + 4 * n);
        }
        for (int k = tid; k < snb[l]; k += 1024) bn.bin_start[boff + k] = moff + lbst[l][k];
        bigoff += tot[l][2];
        n_bins += snb[l];
        n_mem += tot[l][1];
      }
    }
  }
  if (tid == 0) bn.bin_start[n_bins] = n_mem, bn.hdr[0] = 0, bn.hdr[1] = n_bins;
  __syncthreads();
  // ---- hand-out order: heavy bins first ----
  for (int b = tid; b < n_bins; b += 1024) atomicAdd(&hist[63 - min(63, key[b] >> 8)], 1);
  __syncthreads();
  if (warp == 0) {
    const int t0 = hist[2 * lane], t1 = hist[2 * lane + 1];
    int x = t0 + t1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    cursor[2 * lane] = x - t0 - t1;
    cursor[2 * lane + 1] = x - t1;
  }
  __syncthreads();
  for (int b = tid; b < n_bins; b += 1024) bn.order[atomicAdd(&cursor[63 - min(63, key[b] >> 8)], 1)] = b;
}

__global__ void __launch_bounds__(FZ_T, 1) embed_fused_kernel(md_graphs_t G, md_state_t S, md_weights_t W, md_work_t K,
                                                             unsigned char* ws, FusedBins bn, int dbg) {
  // NOTE: no pointer arithmetic on the base -- the compiler must keep seeing the shared state space, otherwise every
  // access below turns into a generic LD/ST (measured: 4x slower gather).  Alignment is asserted instead.
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  FusedSmem& sm = *reinterpret_cast<FusedSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
  const int L = G.n_layers, St = G.stride, B = G.n_graphs;
  const int tpg = St / TILE_M, tiles_per_layer = B * tpg;
  if (tid == 0) {
    mbar_init(&sm.bar_fullH, FZ_GW);
    mbar_init(&sm.bar_fullP, FZ_GW);
    mbar_init(&sm.bar_emptyH, 1);
    sm.long_ctr = 0;
    mbar_init(&sm.bar_emptyP, 1);
    mbar_init(&sm.bar_w, 1);
    mbar_init(&sm.bar_h, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(&sm.tmem, 512);
  // H1 mode (unit cost; same identity as md_agg.cu's agg_deg_kernel<1>): the features are [x, x], x = alive degree / max,
  // so the first embedding is the SAME vector c = relu(w0 + w1) / |relu(w0 + w1)| for every node with an alive neighbour,
  // the first neighbour sum is deg(v) c and the embedding after iteration 0 is H1(d) = normalise(relu(d a + b)) with
  // a = c @ W13, b = c @ W23: iteration 0 (a third of the gather / MMA / epilogue work) is skipped and iteration 1 starts
  // from rows filled with that formula.  md_debug_set_flags bit 18 switches it off.
  const bool h1_mode = G.node_w[0] == nullptr && !(dbg & (1 << 18));
  const int it0 = h1_mode ? 1 : 0;
  if (h1_mode) {
    if (warp == 0) {
      const float ca = fmaxf(W.w_n2l[lane] + W.w_n2l[D + lane], 0.f), cb = fmaxf(W.w_n2l[32 + lane] + W.w_n2l[D + 32 + lane], 0.f);
      const float den = fmaxf(sqrtf(warp_sum(ca * ca + cb * cb)), 1e-12f);
      sm.c0[lane] = ca / den, sm.c0[32 + lane] = cb / den;
    }
    __syncthreads();
    if (tid < D) {
      float a = 0.f, b = 0.f;
      for (int k = 0; k < D; ++k) {
        a = fmaf(sm.c0[k], W.w13[k * D + tid], a);
        b = fmaf(sm.c0[k], W.w23[k * D + tid], b);
      }
      sm.wn2l[tid] = a, sm.wn2l[D + tid] = b;
    }
  } else if (tid < 2 * D) sm.wn2l[tid] = W.w_n2l[tid];
  for (int i = tid; i < (FZ_IDX_CAP + 64) / 8; i += FZ_T) reinterpret_cast<uint4*>(sm.idx)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (tid < 32) sm.zrow[tid] = 0.f;  // every entry is always a valid J (< 8192)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = sm.tmem;
  unsigned char* const spill = ws + FZ_WS_SPILL + (size_t)blockIdx.x * FZ_SPILL_B;
  const unsigned char* const wimg = ws;
  int* const counter = bn.hdr;
  const bool degree_cost = G.node_w[0] != nullptr;
  const bool legacy = (dbg & (1 << 19)) != 0;  // every unit alone and on the single-unit code path (A/B reference)

#ifdef MD_PROFILE
  const long long t_start = clock64();
#endif
  uint32_t t_dbg = 0, c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0, c5 = 0, c6 = 0, c7 = 0, c8 = 0, c9 = 0;  // phase clocks (FZ_CLK)
  uint32_t d0 = 0, d1 = 0, d2 = 0, d3 = 0, d4 = 0, d5 = 0;  // gather sub-phases
  int n_units_done = 0;
  uint32_t uses = 0;    // staging-tile hand-offs so far (same sequence in the gather warps and the MMA warp)
  bool w_ready = false; // this warp has seen the weight images arrive
  int long_base = 0;    // first long-row ticket of the next first tile (the same in every thread)
  uint32_t hloads = 0;  // resident-half reloads consumed (gather warps)
  if (tid == 0) {
    mbar_arrive_expect_tx(&sm.bar_w, 2 * FZ_WPASS_B);
    bulk_g2s(sm.Wt, wimg, 2 * FZ_WPASS_B, &sm.bar_w);
  }

  for (;;) {
    if (tid == 0) {
      const int slot = atomicAdd(counter, 1);
      sm.unit = slot < bn.hdr[1] ? bn.order[slot] : -1;
    }
    __syncthreads();
    const int bin = sm.unit;
    if (bin < 0) break;
    if (warp == 0) {  // the bin's members: graph, rows, entries; prefix sums over the members
      const int m0 = bn.bin_start[bin], nm = min(bn.bin_start[bin + 1] - m0, FZ_MAXM);
      const int lay = bn.member[m0] % L;
      int g_i = 0, n_i = 0, e_i = 0, e0_i = 0;
      if (lane < nm) {
        g_i = bn.member[m0 + lane] / L;
        n_i = max(S.n_act[g_i], 0);
        e0_i = S.aptr[lay][g_i * St];
        e_i = S.aptr[lay][g_i * St + n_i] - e0_i;
      }
      int rs = n_i, es = e_i;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, rs, o), z = __shfl_up_sync(0xffffffffu, es, o);
        if (lane >= o) rs += y, es += z;
      }
      if (lane < nm) {
        sm.mg[lane] = (uint16_t)g_i;
        sm.mrow[lane] = (uint16_t)(rs - n_i);
        sm.meo[lane] = (uint32_t)(es - e_i);
        sm.me0[lane] = e0_i;
      }
      if (lane == nm - 1) {
        sm.mrow[nm] = (uint16_t)rs;
        sm.meo[nm] = (uint32_t)es;
        sm.nm = nm, sm.nrows = rs, sm.nent = es, sm.layer = lay;
        // a unit beyond the packable size runs alone on the single-unit path (column sums from the gather's own rows,
        // node ids translated through shared memory); everything else takes the member-wise path, alone or not
        sm.packed = (legacy || (nm == 1 && (n_i > FZ_SMALL || e_i > FZ_SMALL_E))) ? 0 : 1;
      }
    }
    __syncthreads();
    const int l = sm.layer, n = sm.nrows, nm = sm.nm;
    const bool packed = sm.packed != 0;
    if (n <= 0 || n > FZ_ROWS) continue;
    const int ntile = (n + TILE_M - 1) / TILE_M;
    const int g = sm.mg[0], base = g * St;  // (single-unit path; the packed path reads the member tables)

    // ================================== gather / epilogue warps ==================================
    FZ_CLK(c7);
    ++n_units_done;
    const int gtid = tid;  // 0..FZ_GT-1
    const int gi = lane >> 3, q = lane & 7;
    // tile row of lane group gi in sweep k (k = 0..3): the warp's four rows are neighbours in the degree order (equal length)
    // with swizzle phases {b, b + 1, b + 4, b + 5}: their 8-byte staging stores then cover all 32 banks (two of the rows use
    // chunks 0..3, two use chunks 4..7) instead of piling four rows onto the same 16
    const int rt_w = 8 * (warp >> 1) + 2 * (warp & 1) + 4 * (gi & 1) + (gi >> 1);
    const int e0 = sm.me0[0];
    const int nnz = sm.nent;
    const int32_t* __restrict__ aptr_l = S.aptr[l];
    const int32_t* __restrict__ acol_l = S.acol[l];
    const int32_t* __restrict__ acol_g = acol_l + e0;  // single-unit path
    uint16_t* const wpos_l = bn.wpos + (size_t)l * G.n_total;
    // member of a compact index / of an entry of the bin (binary search over the prefix tables; <= 5 steps)
    auto member_of_row = [&](int p) -> int {
      int lo = 0, hi = nm;
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if ((int)sm.mrow[mid] <= p) lo = mid; else hi = mid;
      }
      return lo;
    };
    auto member_of_entry = [&](int e) -> int {
      int lo = 0, hi = nm;
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if ((int)sm.meo[mid] <= e) lo = mid; else hi = mid;
      }
      return lo;
    };

    if (gtid == 0) sm.cnt0 = 0;  // (ordered before the fill's atomics by the barriers of the sort below)
    // ---- row order: descending alive degree (buckets 31+, 30, ..., 1), stable inside a bucket => deterministic ----
    {
      uint16_t* cnt = sm.idx;  // [32 chunks of 32 rows][32 buckets], reused by the neighbour cache afterwards
      for (int i = gtid; i < 1024; i += FZ_GT) cnt[i] = 0;
      gw_sync();
      int b_[FZ_RPT], rank_[FZ_RPT], beg_[FZ_RPT], d_[FZ_RPT], v_[FZ_RPT];
#pragma unroll
      for (int k = 0; k < FZ_RPT; ++k) {
        const int c = warp + FZ_GW * k, p = 32 * c + lane;
        const bool valid = p < n;
        int a0 = 0, a1 = 0;
        if (valid) {
          const int mi = nm > 1 ? member_of_row(p) : 0;
          const int bs = (int)sm.mg[mi] * St, pl = p - (int)sm.mrow[mi];
          const int eo = (int)sm.meo[mi] - sm.me0[mi];
          a0 = aptr_l[bs + pl] + eo;
          a1 = aptr_l[bs + pl + 1] + eo;
          v_[k] = (S.act_list[bs + pl] - bs) | (mi << 10);
        }
        d_[k] = a1 - a0;
        beg_[k] = a0;
        b_[k] = valid ? min(d_[k], 31) : 32;
        const unsigned m = __match_any_sync(0xffffffffu, b_[k]);
        rank_[k] = __popc(m & ((1u << lane) - 1u));
        if (valid && rank_[k] == 0) cnt[c * 32 + b_[k]] = (uint16_t)__popc(m);
      }
      gw_sync();
      if (warp == 0) {  // lane = bucket: exclusive scan over chunks, then suffix scan over buckets (descending order)
        int running = 0;
        for (int c = 0; c < 32; ++c) {
          const int t = cnt[c * 32 + lane];
          cnt[c * 32 + lane] = (uint16_t)running;
          running += t;
        }
        int x = running;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int y = __shfl_down_sync(0xffffffffu, x, o);
          if (lane + o < 32) x += y;
        }
        sm.bbase[lane] = (uint16_t)(x - running);
      }
      gw_sync();
#pragma unroll
      for (int k = 0; k < FZ_RPT; ++k) {
        const int c = warp + FZ_GW * k;
        if (b_[k] < 32) {
          const int r = sm.bbase[b_[k]] + cnt[c * 32 + b_[k]] + rank_[k];
          if (packed) {
            const uint16_t Jr = (uint16_t)((r << 3) | (r & 7));
            sm.pos[32 * c + lane] = Jr;
            wpos_l[(int)sm.mg[v_[k] >> 10] * St + (v_[k] & 1023)] = Jr;
          } else {
            sm.pos[v_[k]] = (uint16_t)r;
          }
          sm.rowv[r] = (uint16_t)v_[k];
          sm.rbe[r] = (uint32_t)beg_[k] | ((uint32_t)(beg_[k] + d_[k]) << 16);
        }
      }
      gw_sync();
    }
    // neighbour entries are stored as J = 8*j | (j & 7) (j = row of the neighbour): the 16-byte chunk `q` of row j then
    // sits at byte ((J ^ q) << 4) of the resident buffer (SWIZZLE_128B: chunk ^= row & 7)
    {
      const int ncache = min(nnz, FZ_IDX_CAP);
      if (!packed) {
        for (int e = gtid; e < ncache; e += FZ_GT) {
          const uint32_t j = sm.pos[acol_g[e] - base];
          sm.idx[e] = (uint16_t)((j << 3) | (j & 7));
        }
      } else {  // node slot -> row code through the global map this CTA has just written (ordered by the barrier above)
        for (int e = gtid; e < ncache; e += FZ_GT) {
          const int mi = nm > 1 ? member_of_entry(e) : 0;
          const int src = acol_l[sm.me0[mi] + (e - (int)sm.meo[mi])];
          sm.idx[e] = __ldcg(wpos_l + src);
        }
      }
    }
    // (visibility of idx: the barrier after the first resident-half fill below)
    const int n_long = sm.bbase[FZ_LONG];  // rows whose bucket exceeds FZ_LONG, i.e. alive degree > FZ_LONG
    const uint32_t hs_a = smem_u32(sm.Hs), idx_a = smem_u32(sm.idx);
    const uint32_t Jz = (smem_u32(sm.zrow) - hs_a) >> 4;  // "row" of zeros (any chunk of it)
    auto nbrJ = [&](int e) -> uint32_t {  // entries beyond the cache (units with > FZ_IDX_CAP alive entries) come from L2
      if (e < FZ_IDX_CAP) return lds16(idx_a + 2 * e);
      const uint32_t j = sm.pos[acol_g[e] - base];
      return (j << 3) | (j & 7);
    };
    auto ld_J = [&](uint32_t J) -> F4P { return lds128p(hs_a + ((J ^ (uint32_t)q) << 4)); };
    // every MMA issued so far has completed (both groups of the last tile; `uses` tiles handed over so far)
    auto mma_drain = [&]() {
      if (uses > 0) {
        mbar_wait(&sm.bar_emptyH, (uses - 1) & 1);
        mbar_wait(&sm.bar_emptyP, (uses - 1) & 1);
      }
    };
    // MMA group g (0: own rows x W23, 1: neighbour sums x W13) of tile t: a.w = a_hi.w_hi + a_hi.w_lo + a_lo.w_hi with every
    // factor in fp16 (products exact in the fp32 accumulator, the dropped a_lo.w_lo is 2^-22 of the term): 6 kind::f16 MMAs
    // (K = 16) over the staging tile [hi (32) | lo (32)] and the pass's weight tile [W_hi | W_lo]
    auto issue_group = [&](int grp, int t, int pass) {
      if (!w_ready) {
        mbar_wait(&sm.bar_w, 0);
        w_ready = true;
      }
      tc_fence_after();
      if (elect_one()) {
        const uint32_t idb = idesc(FMT_F16, 128, 64);
        const uint64_t da = desc_k_sw128(smem_u32(grp ? sm.P : sm.LO), 0);
        const uint64_t db = desc_k_sw128(smem_u32(sm.Wt) + pass * FZ_WPASS_B + (grp ? 0u : 8192u), 0);
        const uint32_t d = tbase + 64 * t;
#pragma unroll
        for (int k = 0; k < 2; ++k) {  // 32 bytes (16 halves) per K step; hi halves at bytes 0..63, lo at 64..127
          const uint64_t ko = 2 * k;
          mma_f16(d, da + ko, db + ko, idb, (grp | pass | k) != 0);  // the own-row group comes first: it opens the accumulator
          mma_f16(d, da + ko, db + 4 + ko, idb, 1);
          mma_f16(d, da + 4 + ko, db + ko, idb, 1);
        }
        mma_commit(grp ? &sm.bar_emptyP : &sm.bar_emptyH);
      }
      __syncwarp();
    };

    FZ_CLK(c0);
    float inv0[FZ_RPT] = {};
    for (int it = it0; it < MD_BP_ITERS; ++it) {
      for (int pass = 0; pass < 2; ++pass) {
        // ---------------- resident column half ----------------
        if (it == it0) {
          mma_drain();  // (every earlier MMA has completed)
          float xf[FZ_RPT], dmx[FZ_RPT];  // the thread's (up to) FZ_RPT rows: fetch the features first, the loads are independent
          // (inv0[]: the rows' inverse norms, formed in pass 0 and reused in pass 1 -- declared outside the pass loop)
#pragma unroll
          for (int k = 0; k < FZ_RPT; ++k) {
            const int r = gtid + FZ_GT * k;
            xf[k] = 0.f, dmx[k] = 1.f;
            if (r < n) {
              const int rv = sm.rowv[r], gg = sm.mg[rv >> 10];
              const int v = gg * St + (rv & 1023);
              xf[k] = degree_cost ? G.node_w[l][v] : (float)S.deg[l][v];
              if (!degree_cost && !h1_mode) dmx[k] = (float)S.max_deg[l * B + gg];
            }
          }
          if (h1_mode && pass == 0 && !packed) {  // iteration 0's column sums are (rows with an alive neighbour) x c: count them
            int c1n = 0;
#pragma unroll
            for (int k = 0; k < FZ_RPT; ++k) c1n += xf[k] > 0.f ? 1 : 0;
            c1n = warp_sum_i(c1n);
            if (lane == 0 && c1n) atomicAdd(&sm.cnt0, c1n);
          }
#pragma unroll
          for (int k = 0; k < FZ_RPT; ++k) {
            const int r = gtid + FZ_GT * k;
            if (r >= ntile * TILE_M) break;
            float x0 = 0.f, x1 = 0.f;
            if (r < n) {
              if (degree_cost) {
                x0 = xf[k];
                x1 = 1.0f;
              } else if (h1_mode) {  // H1(d) = normalise(relu(d a + b)); d = 0: no neighbour, first embedding 0, H1 = 0
                x0 = xf[k];
                x1 = xf[k] > 0.f ? 1.0f : 0.f;
              } else {
                x0 = xf[k] / dmx[k];  // degree / degree_max, both exact small integers in f32
                x1 = x0;
              }
            }
            const float4* w0v = reinterpret_cast<const float4*>(sm.wn2l);
            const float4* w1v = reinterpret_cast<const float4*>(sm.wn2l + D);
            if (pass == 0) {
              float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;  // four independent chains
#pragma unroll
              for (int c4 = 0; c4 < D / 4; ++c4) {
                const float4 w0 = w0v[c4], w1 = w1v[c4];
                const float a0 = fmaxf(fmaf(x1, w1.x, x0 * w0.x), 0.f), a1 = fmaxf(fmaf(x1, w1.y, x0 * w0.y), 0.f);
                const float a2 = fmaxf(fmaf(x1, w1.z, x0 * w0.z), 0.f), a3 = fmaxf(fmaf(x1, w1.w, x0 * w0.w), 0.f);
                s0 = fmaf(a0, a0, s0);
                s1 = fmaf(a1, a1, s1);
                s2 = fmaf(a2, a2, s2);
                s3 = fmaf(a3, a3, s3);
              }
              inv0[k] = __frcp_rn(fmaxf(sqrtf((s0 + s1) + (s2 + s3)), 1e-12f));
            }
            const float inv = inv0[k];
            unsigned char* rowp = sm.Hs + (uint32_t)r * 128u;
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) {
              const float4 w0 = w0v[8 * pass + ch], w1 = w1v[8 * pass + ch];
              float4 o;
              o.x = fmaxf(fmaf(x1, w1.x, x0 * w0.x), 0.f) * inv;
              o.y = fmaxf(fmaf(x1, w1.y, x0 * w0.y), 0.f) * inv;
              o.z = fmaxf(fmaf(x1, w1.z, x0 * w0.z), 0.f) * inv;
              o.w = fmaxf(fmaf(x1, w1.w, x0 * w0.w), 0.f) * inv;
              *reinterpret_cast<float4*>(rowp + ((ch ^ (r & 7)) << 4)) = o;
            }
          }
          fence_proxy_async_smem();
          gw_sync();
          FZ_CLK(c1);
        } else if (pass == 1) {
          mma_drain();  // pass 0 of this iteration: all gathers and MMAs done
          if (gtid == 0) {
            mbar_arrive_expect_tx(&sm.bar_h, ntile * FZ_TILE_B);
            bulk_g2s(sm.Hs, spill, ntile * FZ_TILE_B, &sm.bar_h);
          }
          mbar_wait(&sm.bar_h, hloads & 1);
          ++hloads;
          FZ_CLK(c2);
        }  // it > 0, pass 0: written by the epilogue below

        // ---------------- tiles: gather -> staging -> MMA ----------------
        // Tile row 4 FZ_GW k + rt_w (k < FZ_RPG): a lane group walks FZ_RPG rows at once, one neighbour of each per step, several
        // steps unrolled (8 independent row loads in flight per lane; the rows are 4 FZ_GW apart in the degree-sorted order, so
        // they end together); the four groups of a warp take neighbouring rows (equal degree; rt_w: swizzle phases chosen so
        // that the half-warps of an 8-byte staging store fall into different bank halves).  Per tile: own rows staged and their
        // MMA group issued first, then the gather, then the neighbour sums staged and the second MMA group.  Rows longer than
        // FZ_LONG (the first n_long rows of the bin) are summed afterwards by whole warps.
        float4 csum = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int t = 0; t < ntile; ++t) {
          const int nl_t = min(max(n_long - TILE_M * t, 0), TILE_M);
          uint32_t ia[FZ_RPG];
          int m[FZ_RPG];
          F4P a4[FZ_RPG];
          int cnt = 0;
          bool all_cached = true;
#pragma unroll
          for (int k = 0; k < FZ_RPG; ++k) {
            const int rt = 4 * FZ_GW * k + rt_w, r = TILE_M * t + rt;
            uint32_t be = 0;
            if (r < n && rt >= nl_t) be = sm.rbe[r];
            ia[k] = idx_a + 2 * (be & 0xffffu);
            m[k] = (int)(be >> 16) - (int)(be & 0xffffu);
            all_cached &= (be >> 16) <= (uint32_t)FZ_IDX_CAP;
            cnt = max(cnt, m[k]);
            a4[k].lo = 0ull, a4[k].hi = 0ull;
          }
          // ---- the rows' own embeddings first (every row of the tile, long or not): they need no gathering, and their half
          // of the tile's MMAs then runs under the gather
          {
            float4 own4[FZ_RPG];
#pragma unroll
            for (int k = 0; k < FZ_RPG; ++k) {
              const int r = TILE_M * t + 4 * FZ_GW * k + rt_w;
              own4[k] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (r < n) own4[k] = unpack4(ld_J((uint32_t)((r << 3) | (r & 7))));
            }
            if (!packed) {
#pragma unroll
              for (int k = 0; k < FZ_RPG; ++k) add4(csum, own4[k]);
            }
            if (uses > 0) mbar_wait(&sm.bar_emptyH, (uses - 1) & 1);  // own-row staging free (previous tile's own-row MMAs complete)
#pragma unroll
            for (int k = 0; k < FZ_RPG; ++k) stage_f16(sm.LO, 4 * FZ_GW * k + rt_w, q, own4[k]);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.bar_fullH);
            if (warp == (int)(uses & (FZ_GW - 1))) {  // this tile's issuer of the own-row group (the warps take turns)
              mbar_wait(&sm.bar_fullH, uses & 1);
              issue_group(0, t, pass);
            }
          }
          FZ_CLK(d0);
          if (all_cached) {
            // branch-free: every step loads one entry of each of the four rows (a finished row re-reads entries just past
            // its end -- always inside the cache array) and a finished row is pointed at the all-zero row, so the four
            // load chains overlap and the adds need no predicate
#pragma unroll 8 / FZ_RPG
            for (int i = 0; i < cnt; ++i) {
              uint32_t J[FZ_RPG];
              F4P v[FZ_RPG];
#pragma unroll
              for (int k = 0; k < FZ_RPG; ++k) {
                const uint32_t Jl = lds16(ia[k] + 2 * i);
                J[k] = i < m[k] ? Jl : Jz;
              }
#pragma unroll
              for (int k = 0; k < FZ_RPG; ++k) v[k] = ld_J(J[k]);
#pragma unroll
              for (int k = 0; k < FZ_RPG; ++k) add4p(a4[k], v[k]);
            }
          } else {  // units with more alive entries than the cache holds
            for (int i = 0; i < cnt; ++i) {
#pragma unroll
              for (int k = 0; k < FZ_RPG; ++k)
                if (i < m[k]) add4p(a4[k], ld_J(nbrJ((int)((ia[k] - idx_a) >> 1) + i)));
            }
          }
          FZ_CLK(d1);
          if (uses > 0) {
            mbar_wait(&sm.bar_emptyP, (uses - 1) & 1);  // P staging free again (the neighbour-sum MMAs of the previous tile complete)
            FZ_CLK(c4);
          }
#pragma unroll
          for (int k = 0; k < FZ_RPG; ++k) {
            const int rt = 4 * FZ_GW * k + rt_w;
            if (rt >= nl_t) stage_f16(sm.P, rt, q, unpack4(a4[k]));
          }
          FZ_CLK(d3);
          // long rows of this tile: one warp per row, lane group gi takes every 4th neighbour.  In the first tile (where all
          // of them are unless more than 128 rows are long) the warps TAKE rows from a counter as they become free -- a static
          // deal left the warp with the largest hubs behind at every tile -- which changes who sums a row, not how: tickets
          // count on from long_base (every warp ends a tile with one ticket beyond its rows; two uses are always separated
          // by the barrier at the end of a pass).
          auto long_row = [&](int rt) {
            const int r = TILE_M * t + rt;
            const uint32_t be = sm.rbe[r];
            const int en = (int)(be >> 16);
            F4P ap = {0ull, 0ull};
            int e = (int)(be & 0xffffu) + gi;
            if (en <= FZ_IDX_CAP) {
              for (; e + 12 < en; e += 16) {  // four independent loads in flight per lane group
                const uint32_t ea = idx_a + 2 * e;
                const uint32_t J0 = lds16(ea), J1 = lds16(ea + 8), J2 = lds16(ea + 16), J3 = lds16(ea + 24);
                const F4P v0 = ld_J(J0), v1 = ld_J(J1), v2 = ld_J(J2), v3 = ld_J(J3);
                add4p(ap, v0);
                add4p(ap, v1);
                add4p(ap, v2);
                add4p(ap, v3);
              }
              for (; e < en; e += 4) add4p(ap, ld_J(lds16(idx_a + 2 * e)));
            } else {
              for (; e < en; e += 4) add4p(ap, ld_J(nbrJ(e)));
            }
            float4 a = unpack4(ap);
            // fixed combination order: (g0 + g1) + (g2 + g3)
            a.x += __shfl_xor_sync(0xffffffffu, a.x, 8);
            a.y += __shfl_xor_sync(0xffffffffu, a.y, 8);
            a.z += __shfl_xor_sync(0xffffffffu, a.z, 8);
            a.w += __shfl_xor_sync(0xffffffffu, a.w, 8);
            a.x += __shfl_xor_sync(0xffffffffu, a.x, 16);
            a.y += __shfl_xor_sync(0xffffffffu, a.y, 16);
            a.z += __shfl_xor_sync(0xffffffffu, a.z, 16);
            a.w += __shfl_xor_sync(0xffffffffu, a.w, 16);
            if (gi == 0) stage_f16(sm.P, rt, q, a);
          };
          if (nl_t > 0) {
            if (t == 0 && !(dbg & (1 << 20))) {
              for (;;) {
                int tk = 0;
                if (lane == 0) tk = atomicAdd(&sm.long_ctr, 1);
                tk = __shfl_sync(0xffffffffu, tk, 0) - long_base;
                if (tk >= nl_t) break;
                long_row(tk);
              }
              long_base += nl_t + FZ_GW;
            } else {
              for (int rt = warp; rt < nl_t; rt += FZ_GW) long_row(rt);
            }
          }
          FZ_CLK(d4);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&sm.bar_fullP);
          if (warp == (int)((uses + FZ_GW / 2) & (FZ_GW - 1))) {  // this tile's issuer of the neighbour-sum group
            mbar_wait(&sm.bar_fullP, uses & 1);
            issue_group(1, t, pass);
          }
          ++uses;
          FZ_CLK(d5);
        }
        // ---------------- end of pass ----------------
        FZ_CLK(c3);
        // column sums of this half of the iteration's input embeddings (spmm subgsum)
        if (packed) {
          // member-wise, straight from the resident half, in the member's own (ascending node) order: one warp per
          // (member, 128 rows) piece, lane = column; independent of the bin's composition.  Runs while the last MMAs of the
          // pass drain (it only reads the resident half).
          const uint32_t pos_a = smem_u32(sm.pos), rbe_a = smem_u32(sm.rbe);
          const uint32_t laneoff = (uint32_t)(((lane >> 2) << 4) | ((lane & 3) << 2));
          int item = 0;
          for (int mi = 0; mi < nm; ++mi) {
            const int p0 = sm.mrow[mi], p1 = sm.mrow[mi + 1];
            for (int c0p = p0, ch = 0; c0p < p1; c0p += TILE_M, ++ch, ++item) {
              if ((item & (FZ_GW - 1)) != warp) continue;
              const int pe = min(c0p + TILE_M, p1);
              float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
              int p = c0p;
              for (; p + 3 < pe; p += 4) {
                const uint32_t J0 = lds16(pos_a + 2 * p), J1 = lds16(pos_a + 2 * p + 2), J2 = lds16(pos_a + 2 * p + 4),
                               J3 = lds16(pos_a + 2 * p + 6);
                float v0, v1, v2, v3;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v0) : "r"(hs_a + ((J0 << 4) ^ laneoff)));
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v1) : "r"(hs_a + ((J1 << 4) ^ laneoff)));
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v2) : "r"(hs_a + ((J2 << 4) ^ laneoff)));
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v3) : "r"(hs_a + ((J3 << 4) ^ laneoff)));
                s0 += v0, s1 += v1, s2 += v2, s3 += v3;
              }
              for (; p < pe; ++p) {
                float v0;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v0) : "r"(hs_a + ((lds16(pos_a + 2 * p) << 4) ^ laneoff)));
                s0 += v0;
              }
              const size_t slot = (size_t)sm.mg[mi] * tpg + ch;
              K.psum[(((size_t)it * L + l) * tiles_per_layer + slot) * D + 32 * pass + lane] = (s0 + s1) + (s2 + s3);
              if (h1_mode && it == it0) {  // the skipped iteration's column sums: (rows with an alive neighbour) x c
                int c1n = 0;
                for (int pp = c0p + lane; pp < pe; pp += 32) {
                  const uint32_t J = lds16(pos_a + 2 * pp);
                  uint32_t be;
                  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(be) : "r"(rbe_a + 4 * (J >> 3)));
                  c1n += (be >> 16) > (be & 0xffffu) ? 1 : 0;
                }
                c1n = warp_sum_i(c1n);
                K.psum[(((size_t)0 * L + l) * tiles_per_layer + slot) * D + 32 * pass + lane] = (float)c1n * sm.c0[32 * pass + lane];
              }
            }
          }
        }
        mma_drain();  // all MMAs of the pass (hence all gathers) are complete
        if (packed) gw_sync();  // ... and every warp's column sums: the resident half may be overwritten from here on
        if (!packed) {  // single unit: the sums of the rows every thread has walked; groups, then warps, fixed order
          csum.x += __shfl_xor_sync(0xffffffffu, csum.x, 8);
          csum.y += __shfl_xor_sync(0xffffffffu, csum.y, 8);
          csum.z += __shfl_xor_sync(0xffffffffu, csum.z, 8);
          csum.w += __shfl_xor_sync(0xffffffffu, csum.w, 8);
          csum.x += __shfl_xor_sync(0xffffffffu, csum.x, 16);
          csum.y += __shfl_xor_sync(0xffffffffu, csum.y, 16);
          csum.z += __shfl_xor_sync(0xffffffffu, csum.z, 16);
          csum.w += __shfl_xor_sync(0xffffffffu, csum.w, 16);
          if (lane < 8) *reinterpret_cast<float4*>(&sm.red[warp][4 * q]) = csum;
          gw_sync();
          if (warp == 0) {
            float s = 0.f;
#pragma unroll
            for (int w2 = 0; w2 < FZ_GW; ++w2) s += sm.red[w2][lane];
            float* ps = K.psum + (((size_t)it * L + l) * tiles_per_layer + (size_t)g * tpg) * D + 32 * pass + lane;
            ps[0] = s;
            for (int t = 1; t < ntile; ++t) ps[(size_t)t * D] = 0.f;
            if (h1_mode && it == it0) {  // the skipped iteration's column sums
              float* ps0 = K.psum + (((size_t)0 * L + l) * tiles_per_layer + (size_t)g * tpg) * D + 32 * pass + lane;
              ps0[0] = (float)sm.cnt0 * sm.c0[32 * pass + lane];
              for (int t = 1; t < ntile; ++t) ps0[(size_t)t * D] = 0.f;
            }
          }
        }
        FZ_CLK(c5);
      }

      // ---------------- epilogue: TMEM -> relu -> normalise -> next resident half / scratch / HBM ----------------
      // One thread per row holds the 64 outputs.  Columns 0..31 go straight into the resident buffer.  Everything bound for
      // global memory is first laid out in the (now idle) staging tiles -- P for warps 0..3, LO for warps 4..7 -- because
      // 16-byte stores scattered at a 128-byte stride cost ~1 cycle each on the SM (measured: 10 k cycles per epilogue):
      // the scratch image leaves as one 4 KB bulk copy per warp and tile (a warp's 32 rows are contiguous in the tile, so
      // the warps of a tile need no barrier among themselves), the final embeddings as fully coalesced 128-byte row pieces.
      tc_fence_after();
      if (warp < 8) {
        const int quad = warp >> 2, lg = warp & 3;
        unsigned char* const stg = quad ? sm.LO : sm.P;
        const uint32_t stg_a = smem_u32(stg);
        const bool last = it + 1 == MD_BP_ITERS;
        for (int t = quad; t < ntile; t += 2) {
          const uint32_t taddr = tbase + ((uint32_t)(32 * lg) << 16) + 64 * t;
          float x[64];
#pragma unroll
          for (int c0 = 0; c0 < 64; c0 += 16) tmem_ld16(taddr + c0, *reinterpret_cast<float(*)[16]>(&x[c0]));
          tmem_ld_wait();
          float ss0 = 0.f, ss1 = 0.f, ss2 = 0.f, ss3 = 0.f;  // four chains
#pragma unroll
          for (int c = 0; c < 64; c += 4) {
            x[c] = fmaxf(x[c], 0.f), x[c + 1] = fmaxf(x[c + 1], 0.f), x[c + 2] = fmaxf(x[c + 2], 0.f), x[c + 3] = fmaxf(x[c + 3], 0.f);
            ss0 = fmaf(x[c], x[c], ss0);
            ss1 = fmaf(x[c + 1], x[c + 1], ss1);
            ss2 = fmaf(x[c + 2], x[c + 2], ss2);
            ss3 = fmaf(x[c + 3], x[c + 3], ss3);
          }
          const float inv = __frcp_rn(fmaxf(sqrtf((ss0 + ss1) + (ss2 + ss3)), 1e-12f));
          const int rt = 32 * lg + lane, r = TILE_M * t + rt;
          FZ_CLK(c6);
          if (!last) {
            unsigned char* hp = sm.Hs + (uint32_t)r * 128u;
#pragma unroll
            for (int ch = 0; ch < 8; ++ch)
              *reinterpret_cast<float4*>(hp + ((ch ^ (rt & 7)) << 4)) =
                  make_float4(x[4 * ch] * inv, x[4 * ch + 1] * inv, x[4 * ch + 2] * inv, x[4 * ch + 3] * inv);
            if (t != quad) {  // this warp's previous bulk copy must have read its rows of the staging tile
              if (lane == 0) bulk_wait_read0();
              __syncwarp();
            }
#pragma unroll
            for (int ch = 0; ch < 8; ++ch)
              *reinterpret_cast<float4*>(stg + (uint32_t)rt * 128u + ((ch ^ (rt & 7)) << 4)) =
                  make_float4(x[32 + 4 * ch] * inv, x[33 + 4 * ch] * inv, x[34 + 4 * ch] * inv, x[35 + 4 * ch] * inv);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              bulk_s2g(spill + (size_t)t * FZ_TILE_B + (uint32_t)lg * 4096u, stg + (uint32_t)lg * 4096u, 4096u);
              bulk_commit();
            }
          } else {
            float* const outb = K.h[MD_BP_ITERS & 1][l];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
#pragma unroll
              for (int ch = 0; ch < 8; ++ch)
                *reinterpret_cast<float4*>(stg + (uint32_t)rt * 128u + ((ch ^ (rt & 7)) << 4)) = make_float4(
                    x[32 * h + 4 * ch] * inv, x[32 * h + 4 * ch + 1] * inv, x[32 * h + 4 * ch + 2] * inv, x[32 * h + 4 * ch + 3] * inv);
              __syncwarp();  // (the warp reads back only the 32 rows it has written itself)
              // copy-out: 8 lanes per row, 4 rows per instruction, 128 contiguous bytes per row
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int rr = 32 * lg + 4 * i + gi, rg = TILE_M * t + rr;
                if (rg < n) {
                  const float4 v = lds128(stg_a + (uint32_t)rr * 128u + (uint32_t)((q ^ (rr & 7)) << 4));
                  const int rv = sm.rowv[rg];
                  *reinterpret_cast<float4*>(outb + (size_t)((int)sm.mg[rv >> 10] * St + (rv & 1023)) * D + 32 * h + 4 * q) = v;
                }
              }
              __syncwarp();
            }
          }
          FZ_CLK(c8);
        }
        if (!last && lane == 0) bulk_wait0();  // this warp's pieces of the scratch image are complete (fetched back for pass 1)
      }
      fence_proxy_async_smem();  // the resident half is read by the tensor core (async proxy) next
      tc_fence_before();
      gw_sync();
      FZ_CLK(c9);
    }
  }
#ifdef MD_PROFILE
  if (tid == 0 && blockIdx.x < 160) g_fz_cta[blockIdx.x][0] = (unsigned)(clock64() - t_start), g_fz_cta[blockIdx.x][1] = n_units_done;
  // gather warp 0: [0] setup [1] fill [2] reload [3] gather [4] stage wait [5] pass end [6] epilogue TMEM+math [7] idle
  //                [12] epilogue stores [13] epilogue fences + barrier
  // MMA warp:      [8] W + resident wait [9] full-tile wait [10] MMA issue [11] completion wait + misc
  if (blockIdx.x == 0 && tid == 0) {
    g_fz_dbg[0] = c0, g_fz_dbg[1] = c1, g_fz_dbg[2] = c2, g_fz_dbg[3] = c3, g_fz_dbg[4] = c4;
    g_fz_dbg[5] = c5, g_fz_dbg[6] = c6, g_fz_dbg[7] = c7, g_fz_dbg[12] = c8, g_fz_dbg[13] = c9, g_fz_dbg[15] = n_units_done;
    g_fz_gath[0] = d0, g_fz_gath[1] = d1, g_fz_gath[2] = d2, g_fz_gath[3] = d3, g_fz_gath[4] = d4, g_fz_gath[5] = d5;
  }
#else
  (void)t_dbg, (void)c0, (void)c1, (void)c2, (void)c3, (void)c4, (void)c5, (void)c6, (void)c7, (void)c8, (void)c9;
  (void)d0, (void)d1, (void)d2, (void)d3, (void)d4, (void)d5, (void)n_units_done;
#endif

  // teardown: the MMA warp still has one weight-image copy in flight
  if (tid == 0) mbar_wait(&sm.bar_w, 0);  // (a CTA that got no bin: its weight copy is still in flight)
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 512);
}

}  // namespace md

extern int g_md_debug_flags;

extern "C" int md_debug_read_fused(long long* host, int n) {
  return (int)cudaMemcpyFromSymbol(host, md::g_fz_dbg, sizeof(long long) * (n < 16 ? n : 16));
}

extern "C" int md_debug_read_fused_gather(unsigned* host8) {
  return (int)cudaMemcpyFromSymbol(host8, md::g_fz_gath, sizeof(unsigned) * 8);
}
extern "C" int md_debug_read_fused_ctas(unsigned* host) {
  return (int)cudaMemcpyFromSymbol(host, md::g_fz_cta, sizeof(unsigned) * 320);
}

extern "C" int32_t md_fused_max_stride(void) { return md::FZ_ROWS; }
// row pointers of one unit are kept as uint16 offsets: a graph's CSR entries (all layers) must stay below 65536
extern "C" int32_t md_fused_applies(const md_graphs_t* g) {
  return g && g->stride <= md::FZ_ROWS && g->max_graph_nnz > 0 && g->max_graph_nnz < 65536;
}

static int fused_sm_count() {
  static md::PerDevice<int> nsm_pd;
  int& nsm = nsm_pd.get();
  if (nsm <= 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    if (nsm <= 0) nsm = 148;
  }
  return nsm;
}

static size_t fused_order_off() { return md::FZ_WS_SPILL + (size_t)fused_sm_count() * md::FZ_SPILL_B; }

// scratch behind the spill images: bin header, hand-out order, bin starts, members, node slot -> row code map
static md::FusedBins fused_bins(unsigned char* ws, const md_graphs_t* g) {
  const size_t nu = (size_t)g->n_graphs * (size_t)g->n_layers;
  md::FusedBins b;
  int* p = reinterpret_cast<int*>(ws + fused_order_off());
  b.hdr = p;
  b.order = p + 4;
  b.bin_start = b.order + nu;
  b.member = b.bin_start + nu + 1;
  b.wpos = reinterpret_cast<uint16_t*>(b.member + nu + 1);
  return b;
}

extern "C" int64_t md_fused_ws_bytes(const md_graphs_t* g) {
  if (!g) return 0;
  const size_t nu = (size_t)g->n_graphs * (size_t)g->n_layers;
  return (int64_t)(fused_order_off() + sizeof(int) * (3 * nu + 8) + sizeof(uint16_t) * (size_t)g->n_layers * (size_t)g->n_total + 64);
}

static int g_md_fused_pack = 0;
extern "C" void md_debug_set_fused_pack(int32_t rows) { g_md_fused_pack = rows; }

extern "C" int md_embed_fused(const md_graphs_t* g, const md_state_t* s, const md_weights_t* w, const md_work_t* k,
                              md_stream_t stream) {
  if (!g || !s || !w || !k || !k->fused_ws) return MD_EINVAL;
  if (g->n_layers < 1 || g->n_layers > MD_MAX_LAYERS) return MD_ELAYERS;
  if (g->stride % MD_TILE_M != 0 || g->n_total != g->n_graphs * g->stride) return MD_EINVAL;
  if (!md_fused_applies(g)) return MD_ETOOBIG;
  static md::PerDevice<bool> configured_pd;
  bool& configured = configured_pd.get();
  const int smem = (int)sizeof(md::FusedSmem);
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(md::embed_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  unsigned char* ws = reinterpret_cast<unsigned char*>(k->fused_ws);
  const int n_units = g->n_graphs * g->n_layers;
  const int nsm = fused_sm_count();
  const int grid = nsm < n_units ? nsm : n_units;
  const md::FusedBins bn = fused_bins(ws, g);
  const int pack = (g_md_debug_flags & (1 << 19)) ? -1 : g_md_fused_pack;
  md::fused_prep_kernel<<<2, 1024, 0, (cudaStream_t)stream>>>(*g, *s, *w, ws, bn, n_units, nsm, pack);
  MD_CHECK_LAUNCH();
  md::embed_fused_kernel<<<grid, md::FZ_T, smem, (cudaStream_t)stream>>>(*g, *s, *w, *k, ws, bn, g_md_debug_flags);
  MD_CHECK_LAUNCH();
  return 0;
}
