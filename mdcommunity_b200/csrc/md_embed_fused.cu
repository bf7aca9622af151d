// md_embed_fused: the node-embedding stage of one forward (first embedding + the three message-passing iterations)
// for batches of small graphs (stride <= 1024), one launch, embeddings resident in shared memory.
// Replaces /root/reference/code/MultiDismantler_unit_cost/MultiDismantler_net_graphsage.py:248-310 (node features,
// relu(x @ w_n2l), normalise; 3 x { torch_sparse.spmm(n2nsum), the three dense transforms, relu, normalise,
// torch_sparse.spmm(subgsum) }) -- the same arithmetic md_embed_init + 3 x md_embed_iter perform through HBM/L2.
//
// Why a second formulation: ncu on md_embed_iter showed the neighbour gather pinned at the L2 -> SM path (one 256-byte
// row per alive edge, ~37 B/cycle/SM at best) with the FP32-FMA transform fighting it for issue slots.  For graphs of
// <= 1024 active nodes one graph-layer's embeddings are 256 KB: too large for one SM's shared memory in full, but one
// COLUMN HALF (32 of 64 columns, 128 B per row) is 128 KB.  So a persistent CTA owns one (graph, layer) unit at a time:
//
//   per iteration, two passes p = 0,1 over the column halves:
//     resident  Hs = H[:, 32p:32p+32] in shared memory, compact row order, stored directly in the tcgen05
//               SWIZZLE_128B K-major tile layout (a 128-byte row is exactly one swizzle span), so the very same bytes
//               serve the gather (8 lanes x LDS.128 per neighbour row, conflict-free) and, as MMA operand A, the
//               "own embedding" term of the transform.
//     gather    16 warps; 8 lanes per row; neighbour lists are read from a shared-memory copy of the unit's alive CSR
//               (compact uint16 indices, translated once per unit); sums in CSR order; rows longer than FZ_LONG are
//               spread over the warp's four lane groups.  Per 128-row tile the warps write P = A_adj.H (raw fp32) and
//               the bf16 remainders [P - tf32(P) | H - tf32(H)] into a staging tile.
//     transform one elected thread issues tcgen05.mma: D_tile += P.W13[32p:32p+32] + H.W23[32p:32p+32] as
//               kind::tf32 (operands are truncated to tf32 by the tensor core -- measured, scripts/microbench/umma_probe.cu)
//               with hi and lo planes of the weights, plus one kind::f16 (bf16) pass for the operand remainders:
//               a.w ~= a_hi.w_hi + a_hi.w_lo + a_lo.w   (error ~2^-20 relative; the probe measures 1.3e-6 of max |D|).
//               Accumulators for ALL tiles of the unit (<= 8 x 64 columns = the whole 512-column TMEM) stay in tensor
//               memory across the two passes, which is what makes the column split free: no partial sums are exchanged.
//   epilogue    tcgen05.ld (one thread per row holds the 64 outputs): relu, L2 normalise; columns 0..31 go straight into
//               the resident buffer for the next iteration's pass 0, columns 32..63 into a per-CTA scratch image
//               (L2-resident) that one cp.async.bulk brings back for pass 1.  The last iteration writes H to HBM for
//               md_qhead.  Per-graph column sums of every iteration's input (the subgsum spmm) fall out of the gather.
//
// HBM traffic per unit and forward: alive CSR once + final H once (+ features); everything else stays on chip / in L2.
#include <cuda_bf16.h>
#include "md_common.cuh"
#include "md_umma.cuh"

namespace md {
using namespace umma;

constexpr int FZ_ROWS = 1024;          // largest stride (node slots per graph) the resident path handles
constexpr int FZ_GW = 16;              // gather / epilogue warps
constexpr int FZ_GT = FZ_GW * 32;
constexpr int FZ_T = FZ_GT + 32;       // + the MMA / bulk-copy warp
constexpr int FZ_IDX_CAP = 8192;       // alive directed edges of one unit cached in shared memory
constexpr int FZ_LONG = 24;            // rows with more alive neighbours are summed by the whole warp
constexpr uint32_t FZ_TILE_B = 16384;  // one 128-row x 128-byte tile
constexpr uint32_t FZ_WPASS_B = 40960; // B operands of one pass: W13 hi, W13 lo, W23 hi, W23 lo (tf32), [W13|W23] bf16
constexpr size_t FZ_WS_CTR = 2 * FZ_WPASS_B;
constexpr size_t FZ_WS_SPILL = FZ_WS_CTR + 1024;
constexpr size_t FZ_SPILL_B = 8 * FZ_TILE_B;

struct FusedSmem {
  unsigned char Hs[8 * FZ_TILE_B];  // resident column half of the unit's embeddings (SWIZZLE_128B tiles)
  unsigned char P[FZ_TILE_B];       // staging: neighbour sums of one tile, raw fp32
  unsigned char LO[FZ_TILE_B];      // staging: bf16 remainders, row = [P_lo (32) | H_lo (32)]
  unsigned char Wt[FZ_WPASS_B];     // B operands of the running pass
  uint16_t idx[FZ_IDX_CAP];         // compact neighbour rows of the unit's alive CSR
  uint16_t pos[FZ_ROWS];            // graph-local node id -> compact row
  uint16_t rptr[FZ_ROWS + 8];       // row pointers into idx
  float red[FZ_GW][32];
  float wn2l[2 * D];
  uint64_t bar_full, bar_empty, bar_w, bar_h;
  uint32_t tmem;
  int unit;
};

__device__ __forceinline__ void gw_sync() { asm volatile("bar.sync 1, %0;" ::"n"(FZ_GT) : "memory"); }

__device__ __forceinline__ void add4(float4& a, const float4& b) {
  a.x += b.x;
  a.y += b.y;
  a.z += b.z;
  a.w += b.w;
}
__device__ __forceinline__ uint2 lo_bf16x4(const float4& v) {
  const __nv_bfloat162 a = __floats2bfloat162_rn(v.x - tf32_hi(v.x), v.y - tf32_hi(v.y));
  const __nv_bfloat162 b = __floats2bfloat162_rn(v.z - tf32_hi(v.z), v.w - tf32_hi(v.w));
  return make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
}

// Weight images (both passes) in the exact shared-memory layout + reset of the unit counter.
__global__ void __launch_bounds__(256) fused_prep_kernel(md_weights_t W, unsigned char* ws) {
  if (threadIdx.x == 0) *reinterpret_cast<int*>(ws + FZ_WS_CTR) = 0;
  for (int i = threadIdx.x; i < 2 * 64 * 32; i += 256) {
    const int p = i >> 11, n = (i >> 5) & 63, k = i & 31;
    const float w13 = W.w13[(32 * p + k) * D + n], w23 = W.w23[(32 * p + k) * D + n];
    unsigned char* img = ws + (size_t)p * FZ_WPASS_B;
    const uint32_t off = sw128_off(n, k >> 2) + 4 * (k & 3);
    *reinterpret_cast<float*>(img + off) = tf32_hi(w13);
    *reinterpret_cast<float*>(img + 8192 + off) = w13 - tf32_hi(w13);
    *reinterpret_cast<float*>(img + 16384 + off) = tf32_hi(w23);
    *reinterpret_cast<float*>(img + 24576 + off) = w23 - tf32_hi(w23);
    *reinterpret_cast<__nv_bfloat16*>(img + 32768 + sw128_off(n, k >> 3) + 2 * (k & 7)) = __float2bfloat16_rn(w13);
    *reinterpret_cast<__nv_bfloat16*>(img + 32768 + sw128_off(n, 4 + (k >> 3)) + 2 * (k & 7)) = __float2bfloat16_rn(w23);
  }
}

__global__ void __launch_bounds__(FZ_T, 1) embed_fused_kernel(md_graphs_t G, md_state_t S, md_weights_t W, md_work_t K,
                                                             unsigned char* ws, int n_units) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  FusedSmem& sm = *reinterpret_cast<FusedSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int L = G.n_layers, St = G.stride, B = G.n_graphs;
  const int tpg = St / TILE_M, tiles_per_layer = B * tpg;
  if (tid == 0) {
    mbar_init(&sm.bar_full, FZ_GW);
    mbar_init(&sm.bar_empty, 1);
    mbar_init(&sm.bar_w, 1);
    mbar_init(&sm.bar_h, 1);
    mbar_fence_init();
  }
  if (warp == FZ_GW) tmem_alloc(&sm.tmem, 512);
  if (tid < 2 * D) sm.wn2l[tid] = W.w_n2l[tid];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = sm.tmem;
  unsigned char* const spill = ws + FZ_WS_SPILL + (size_t)blockIdx.x * FZ_SPILL_B;
  const unsigned char* const wimg = ws;
  int* const counter = reinterpret_cast<int*>(ws + FZ_WS_CTR);
  const bool degree_cost = G.node_w[0] != nullptr;

  uint32_t uses = 0;    // staging-tile hand-offs so far (same sequence in the gather warps and the MMA warp)
  uint32_t wloads = 0;  // weight images consumed (MMA warp)
  uint32_t hloads = 0;  // resident-half reloads consumed (gather warps)
  if (warp == FZ_GW && lane == 0) {
    mbar_arrive_expect_tx(&sm.bar_w, FZ_WPASS_B);
    bulk_g2s(sm.Wt, wimg, FZ_WPASS_B, &sm.bar_w);
  }

  for (;;) {
    if (tid == 0) sm.unit = atomicAdd(counter, 1);
    __syncthreads();
    const int unit = sm.unit;
    __syncthreads();
    if (unit >= n_units) break;
    const int g = unit / L, l = unit - g * L;
    const int n = S.n_act[g];
    if (n <= 0) continue;
    const int ntile = (n + TILE_M - 1) / TILE_M;
    const int base = g * St;

    if (warp == FZ_GW) {
      // ============================== MMA / bulk-copy warp ==============================
      if (lane == 0) {
        const uint32_t idt = idesc(FMT_TF32, 128, 64), idb = idesc(FMT_BF16, 128, 64);
        const uint64_t dP = desc_k_sw128(smem_u32(sm.P), 0), dLO = desc_k_sw128(smem_u32(sm.LO), 0);
        const uint64_t dH0 = desc_k_sw128(smem_u32(sm.Hs), 0);
        const uint64_t b13h = desc_k_sw128(smem_u32(sm.Wt), 0), b13l = desc_k_sw128(smem_u32(sm.Wt) + 8192, 0);
        const uint64_t b23h = desc_k_sw128(smem_u32(sm.Wt) + 16384, 0), b23l = desc_k_sw128(smem_u32(sm.Wt) + 24576, 0);
        const uint64_t bbf = desc_k_sw128(smem_u32(sm.Wt) + 32768, 0);
        for (int it = 0; it < MD_BP_ITERS; ++it) {
          for (int pass = 0; pass < 2; ++pass) {
            mbar_wait(&sm.bar_w, wloads & 1);
            ++wloads;
            for (int t = 0; t < ntile; ++t) {
              mbar_wait(&sm.bar_full, uses & 1);
              tc_fence_after();
              const uint32_t d = tbase + 64 * t;
              const uint64_t dH = dH0 + (uint64_t)((t * FZ_TILE_B) >> 4);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint64_t ko = 2 * k;  // 32 bytes per K step, in 16-byte units
                mma_tf32(d, dP + ko, b13h + ko, idt, (pass | k) != 0);
                mma_tf32(d, dP + ko, b13l + ko, idt, 1);
                mma_tf32(d, dH + ko, b23h + ko, idt, 1);
                mma_tf32(d, dH + ko, b23l + ko, idt, 1);
                mma_bf16(d, dLO + ko, bbf + ko, idb, 1);
              }
              mma_commit(&sm.bar_empty);
              ++uses;
            }
            // pass complete -> the weight tile is free: fetch the other pass's image
            mbar_wait(&sm.bar_empty, (uses - 1) & 1);
            mbar_arrive_expect_tx(&sm.bar_w, FZ_WPASS_B);
            bulk_g2s(sm.Wt, wimg + (size_t)(pass ^ 1) * FZ_WPASS_B, FZ_WPASS_B, &sm.bar_w);
          }
        }
      } else {
        uses += 2 * MD_BP_ITERS * ntile;
      }
      uses = __shfl_sync(0xffffffffu, uses, 0);
      wloads = __shfl_sync(0xffffffffu, wloads, 0);
      continue;
    }

    // ================================== gather / epilogue warps ==================================
    const int gtid = tid;  // 0..511
    const int gi = lane >> 3, q = lane & 7;
    const int e0 = S.aptr[l][base];
    const int nnz = S.aptr[l][base + n] - e0;
    const bool cached = nnz <= FZ_IDX_CAP;
    const int32_t* __restrict__ aptr_g = S.aptr[l] + base;
    const int32_t* __restrict__ acol_g = S.acol[l] + e0;
    for (int p = gtid; p < n; p += FZ_GT) sm.pos[S.act_list[base + p] - base] = (uint16_t)p;
    if (cached)
      for (int p = gtid; p <= n; p += FZ_GT) sm.rptr[p] = (uint16_t)(aptr_g[p] - e0);
    gw_sync();
    if (cached)
      for (int e = gtid; e < nnz; e += FZ_GT) sm.idx[e] = sm.pos[acol_g[e] - base];
    // (visibility of idx / rptr: the barrier after the first resident-half fill below)

    auto row_ptr = [&](int r) -> int { return cached ? (int)sm.rptr[r] : aptr_g[r] - e0; };
    auto nbr = [&](int e) -> int { return cached ? (int)sm.idx[e] : (int)sm.pos[acol_g[e] - base]; };
    auto ld_h = [&](int j) -> float4 {
      return *reinterpret_cast<const float4*>(sm.Hs + (uint32_t)j * 128u + (uint32_t)((q ^ (j & 7)) << 4));
    };

    for (int it = 0; it < MD_BP_ITERS; ++it) {
      for (int pass = 0; pass < 2; ++pass) {
        // ---------------- resident column half ----------------
        if (it == 0) {
          if (uses > 0) mbar_wait(&sm.bar_empty, (uses - 1) & 1);  // every earlier MMA (they read Hs) has completed
          const float dmax = (float)S.max_deg[l * B + g];
          for (int r = gtid; r < ntile * TILE_M; r += FZ_GT) {
            float x0 = 0.f, x1 = 0.f;
            if (r < n) {
              const int v = S.act_list[base + r];
              if (degree_cost) {
                x0 = G.node_w[l][v];
                x1 = 1.0f;
              } else {
                x0 = (float)S.deg[l][v] / dmax;
                x1 = x0;
              }
            }
            float ss = 0.f;
#pragma unroll 8
            for (int c = 0; c < D; ++c) {
              const float a = fmaxf(fmaf(x1, sm.wn2l[D + c], x0 * sm.wn2l[c]), 0.f);
              ss = fmaf(a, a, ss);
            }
            const float den = fmaxf(sqrtf(ss), 1e-12f);
            unsigned char* rowp = sm.Hs + (uint32_t)r * 128u;
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) {
              float4 o;
              const int c = 32 * pass + 4 * ch;
              o.x = fmaxf(fmaf(x1, sm.wn2l[D + c], x0 * sm.wn2l[c]), 0.f) / den;
              o.y = fmaxf(fmaf(x1, sm.wn2l[D + c + 1], x0 * sm.wn2l[c + 1]), 0.f) / den;
              o.z = fmaxf(fmaf(x1, sm.wn2l[D + c + 2], x0 * sm.wn2l[c + 2]), 0.f) / den;
              o.w = fmaxf(fmaf(x1, sm.wn2l[D + c + 3], x0 * sm.wn2l[c + 3]), 0.f) / den;
              *reinterpret_cast<float4*>(rowp + ((ch ^ (r & 7)) << 4)) = o;
            }
          }
          fence_proxy_async_smem();
          gw_sync();
        } else if (pass == 1) {
          mbar_wait(&sm.bar_empty, (uses - 1) & 1);  // pass 0 of this iteration: all gathers and MMAs done
          if (gtid == 0) {
            mbar_arrive_expect_tx(&sm.bar_h, ntile * FZ_TILE_B);
            bulk_g2s(sm.Hs, spill, ntile * FZ_TILE_B, &sm.bar_h);
          }
          mbar_wait(&sm.bar_h, hloads & 1);
          ++hloads;
        }  // it > 0, pass 0: written by the epilogue below

        // ---------------- tiles: gather -> staging -> MMA ----------------
        float4 csum = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int t = 0; t < ntile; ++t) {
#pragma unroll 1
          for (int round = 0; round < 2; ++round) {
            const int rt = 8 * warp + 4 * round + gi;
            const int r = TILE_M * t + rt;
            const bool valid = r < n;
            int beg = 0, end = 0;
            if (valid) {
              beg = row_ptr(r);
              end = row_ptr(r + 1);
            }
            const bool is_long = end - beg > FZ_LONG;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            if (!is_long) {
              int e = beg;
              for (; e + 1 < end; e += 2) {
                const int j0 = nbr(e), j1 = nbr(e + 1);
                const float4 v0 = ld_h(j0), v1 = ld_h(j1);
                add4(acc, v0);
                add4(acc, v1);
              }
              if (e < end) add4(acc, ld_h(nbr(e)));
            }
            unsigned longmask = __ballot_sync(0xffffffffu, is_long && q == 0);
            while (longmask) {  // warp-uniform
              const int src = __ffs(longmask) - 1;
              longmask &= longmask - 1;
              const int b = __shfl_sync(0xffffffffu, beg, src), en = __shfl_sync(0xffffffffu, end, src);
              float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
              for (int e = b + gi; e < en; e += 4) add4(a, ld_h(nbr(e)));
              // fixed combination order: (g0 + g1) + (g2 + g3)
              a.x += __shfl_xor_sync(0xffffffffu, a.x, 8);
              a.y += __shfl_xor_sync(0xffffffffu, a.y, 8);
              a.z += __shfl_xor_sync(0xffffffffu, a.z, 8);
              a.w += __shfl_xor_sync(0xffffffffu, a.w, 8);
              a.x += __shfl_xor_sync(0xffffffffu, a.x, 16);
              a.y += __shfl_xor_sync(0xffffffffu, a.y, 16);
              a.z += __shfl_xor_sync(0xffffffffu, a.z, 16);
              a.w += __shfl_xor_sync(0xffffffffu, a.w, 16);
              if (gi == (src >> 3)) acc = a;
            }
            float4 own = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) {
              own = ld_h(r);
              add4(csum, own);
            }
            if (round == 0 && uses > 0) mbar_wait(&sm.bar_empty, (uses - 1) & 1);  // staging tile free again
            const uint32_t ro = (uint32_t)rt * 128u;
            *reinterpret_cast<float4*>(sm.P + ro + ((q ^ (rt & 7)) << 4)) = acc;
            *reinterpret_cast<uint2*>(sm.LO + ro + (((q >> 1) ^ (rt & 7)) << 4) + 8 * (q & 1)) = lo_bf16x4(acc);
            *reinterpret_cast<uint2*>(sm.LO + ro + (((4 + (q >> 1)) ^ (rt & 7)) << 4) + 8 * (q & 1)) = lo_bf16x4(own);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&sm.bar_full);
          ++uses;
        }
        // ---------------- end of pass ----------------
        mbar_wait(&sm.bar_empty, (uses - 1) & 1);  // all MMAs of the pass (hence all gathers) are complete
        // column sums of this half of the iteration's input embeddings (spmm subgsum): groups, then warps, fixed order
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
        }
      }

      // ---------------- epilogue: TMEM -> relu -> normalise -> next resident half / scratch / HBM ----------------
      tc_fence_after();
      {
        const int quad = warp >> 2, lg = warp & 3;
        for (int t = quad; t < ntile; t += 4) {
          const uint32_t taddr = tbase + ((uint32_t)(32 * lg) << 16) + 64 * t;
          float x[64];
#pragma unroll
          for (int c0 = 0; c0 < 64; c0 += 16) tmem_ld16(taddr + c0, *reinterpret_cast<float(*)[16]>(&x[c0]));
          tmem_ld_wait();
          float ss = 0.f;
#pragma unroll
          for (int c = 0; c < 64; ++c) {
            x[c] = fmaxf(x[c], 0.f);
            ss = fmaf(x[c], x[c], ss);
          }
          const float inv = __frcp_rn(fmaxf(sqrtf(ss), 1e-12f));
          const int rt = 32 * lg + lane, r = TILE_M * t + rt;
          if (it + 1 < MD_BP_ITERS) {
            unsigned char* hp = sm.Hs + (uint32_t)r * 128u;
            unsigned char* sp = spill + (size_t)r * 128u;
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) {
              const uint32_t o = (uint32_t)((ch ^ (rt & 7)) << 4);
              *reinterpret_cast<float4*>(hp + o) = make_float4(x[4 * ch] * inv, x[4 * ch + 1] * inv, x[4 * ch + 2] * inv, x[4 * ch + 3] * inv);
              *reinterpret_cast<float4*>(sp + o) =
                  make_float4(x[32 + 4 * ch] * inv, x[33 + 4 * ch] * inv, x[34 + 4 * ch] * inv, x[35 + 4 * ch] * inv);
            }
          } else if (r < n) {
            float* out = K.h[MD_BP_ITERS & 1][l] + (size_t)S.act_list[base + r] * D;
#pragma unroll
            for (int ch = 0; ch < 16; ++ch)
              *reinterpret_cast<float4*>(out + 4 * ch) = make_float4(x[4 * ch] * inv, x[4 * ch + 1] * inv, x[4 * ch + 2] * inv, x[4 * ch + 3] * inv);
          }
        }
      }
      fence_proxy_async_all();  // resident half (shared) and scratch image (global) are read by the async proxy next
      tc_fence_before();
      gw_sync();
    }
  }

  // teardown: the MMA warp still has one weight-image copy in flight
  if (warp == FZ_GW && lane == 0) mbar_wait(&sm.bar_w, wloads & 1);
  tc_fence_before();
  __syncthreads();
  if (warp == FZ_GW) tmem_dealloc(tbase, 512);
}

}  // namespace md

extern int g_md_debug_flags;

extern "C" int32_t md_fused_max_stride(void) { return md::FZ_ROWS; }

static int fused_sm_count() {
  static int nsm = 0;
  if (nsm <= 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    if (nsm <= 0) nsm = 148;
  }
  return nsm;
}

extern "C" int64_t md_fused_ws_bytes(void) {
  return (int64_t)(md::FZ_WS_SPILL + (size_t)fused_sm_count() * md::FZ_SPILL_B);
}

extern "C" int md_embed_fused(const md_graphs_t* g, const md_state_t* s, const md_weights_t* w, const md_work_t* k,
                              md_stream_t stream) {
  if (!g || !s || !w || !k || !k->fused_ws) return MD_EINVAL;
  if (g->n_layers < 1 || g->n_layers > MD_MAX_LAYERS) return MD_ELAYERS;
  if (g->stride % MD_TILE_M != 0 || g->n_total != g->n_graphs * g->stride) return MD_EINVAL;
  if (g->stride > md::FZ_ROWS) return MD_ETOOBIG;
  static bool configured = false;
  const int smem = (int)sizeof(md::FusedSmem) + 1024;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(md::embed_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  unsigned char* ws = reinterpret_cast<unsigned char*>(k->fused_ws);
  const int n_units = g->n_graphs * g->n_layers;
  const int nsm = fused_sm_count();
  const int grid = nsm < n_units ? nsm : n_units;
  md::fused_prep_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(*w, ws);
  MD_CHECK_LAUNCH();
  md::embed_fused_kernel<<<grid, md::FZ_T, smem, (cudaStream_t)stream>>>(*g, *s, *w, *k, ws, n_units);
  MD_CHECK_LAUNCH();
  return 0;
}
