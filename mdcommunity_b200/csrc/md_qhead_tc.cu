// md_qhead on the tensor cores: inter-layer fusion of the node rows + Q head + masked arg-max (same maths and the same
// reference lines as md_qhead.cu: MRGNN/mutil_layer_weight.py:266-313, MultiDismantler_net_graphsage.py:323-393,
// MultiDismantler_torch.py:286-300,799-807), restructured after the FP32-FMA kernel had become the largest share of a
// rollout (ncu: 36 % FMA-pipe utilisation, static tile -> CTA map idle for most of the rollout).
//
//   persistent CTA per SM, 11 warps, dynamic tile queue (live (graph, 128-row) tiles only):
//   warps 8..9  loaders : gather the tile's H_l rows from HBM/L2 (32 row fetches per lane in flight, issued before the wait for
//                         the operand tiles; the next layer's rows are prefetched into L2), store them as raw
//                         fp32 into SWIZZLE_128B K-major operand tiles + the fp16 remainders (x - tf32(x))
//   warp  10    MMA     : T_pre_l = H_l @ trans as the 3-term product (tf32 hi x hi, hi x lo, fp16 remainder pass), one
//                         elected thread, accumulators double-buffered in tensor memory; the operand tile is released to
//                         the loaders as soon as the MMAs have read it, so the next tile's gather overlaps the epilogue
//   warps 0..7  epilogue: two threads per node row (column halves for the tanh, hidden-unit halves for h1): tcgen05.ld, tanh(. + bias), the pairwise (T_k . T_p . lw) gates,
//                         softmax-of-sigmoid mix, normalise, scale by the per-graph s_p, relu(e_p @ h1) . h2, layer-gate mix,
//                         q[v], packed arg-max key.
//   HTC (default): the tanh outputs T are written back IN PLACE over the accumulators (tcgen05.st) and stay in tensor
//   memory for the rest of the tile's epilogue -- every later pass re-reads its 16-column chunks with tcgen05.ld, shared
//   memory carries no T at all (the FP32 variant's Ts[L][64][128] was 64 / 96 KB and 68 % of the LSU pipe) -- and
//   relu's argument e_p @ h1 is a second tensor-core product (operand tiles written by the epilogue).  2 x 64 L accumulator
//   columns + 128 for the hidden layer = 384 (L = 2) / 512 (L = 3) of the 512 TMEM columns.
#include <cuda_fp16.h>
#include "md_common.cuh"
#include "md_umma.cuh"

namespace md {
using namespace umma;

constexpr int QT_T = 352;         // 8 epilogue + 2 loader + 1 MMA warps
constexpr int QT_EPI = 256;
constexpr uint32_t QT_KB = 16384;  // one 128-row x 128-byte operand tile
constexpr int QT_PMAX = 2048;      // batches of up to this many graphs: tiles are cut from the CONCATENATED active lists

// HTC: T kept in tensor memory, relu(e_p @ h1) on the tensor cores
template <int L, bool HTC>
struct QTSmem {
  unsigned char A[2][QT_KB];     // raw fp32 rows of ONE layer's H_l (the layers of a tile pass through in turn), two K blocks
  unsigned char ALO[QT_KB];      // fp16 remainders of the same rows, 64 columns per row
  unsigned char Bt[40960];       // trans^T: tf32 hi (2 K blocks), tf32 lo (2 K blocks), fp16 (64 columns)
  unsigned char E[HTC ? 2 : 1][HTC ? 2 : 1][HTC ? QT_KB : 16];  // HTC: [buffer][K block] X_p = fused rows of one layer (raw fp32);
                                                                // two buffers: layer p + 1 is written while layer p's MMAs run
  unsigned char ELO[HTC ? 2 : 1][HTC ? QT_KB : 16];             // HTC: [buffer] their fp16 remainders
  unsigned char Bh[HTC ? 20480 : 16];               // HTC: h1^T (32 hidden units x 64): tf32 hi (2 x 4 KB), tf32 lo (2 x 4 KB), fp16 (4 KB)
  alignas(16) float h1[HTC ? 4 : D * MD_HIDDEN];
  alignas(16) float bias[D];
  alignas(16) float lw[D];
  alignas(16) float h2[MD_HIDDEN];
  float Ts[HTC ? 1 : L][HTC ? 1 : D][TILE_M];  // FP32 variant: tanh outputs, column-major so that a thread owns one conflict-free column
  float dpart[2][HTC ? L * (L + 1) / 2 : 1][TILE_M];  // HTC: [thread of the row] partial gate dot products over its 32 columns
  float hq[L][TILE_M];           // epilogue: partial hidden sums of the second thread of each row
  float ssq[2][2][HTC ? TILE_M : 1];  // HTC: [layer parity][thread of the row] half of |X_p|^2
  int vrow[4][TILE_M];           // [tile number & 3] global node id of the row, -1 = padding
  int pfx[QT_PMAX + 1];          // compact tiles: exclusive prefix sums of the graphs' active-node counts
  int wtot[12];
  int desc[4][4];                // [tile number & 3] tile, graph
  int next_tile;
  uint64_t a_full, a_empty, d_full[2], d_empty[2], h_full;
  uint32_t tmem;
};

__device__ __forceinline__ float4 qt_lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
// round-to-nearest tf32 (the tensor core truncates fp32 operands: an operand that is already a tf32 value passes unchanged,
// and the remainder x - rn(x) is at most 2^-12 |x| instead of truncation's 2^-11: one more bit for the fp16 / tf32 remainder planes)
__device__ __forceinline__ float tf32_rn(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}
// tanh(x) = 1 - 2 / (e^2x + 1) on the SFU (ex2.approx + rcp.approx): absolute error <= ~3e-7 over the whole range (the exact
// tanhf costs ~25 instructions, 96 of them per thread and tile); saturates correctly (e -> inf: 1, e -> 0: -1)
__device__ __forceinline__ float tanh_fast(float x) { return 1.f - __fdividef(2.f, __expf(2.f * x) + 1.f); }
__device__ __forceinline__ bool qt_elect() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xffffffff;\n\t@px mov.s32 %0, 1;\n\t}" : "+r"(pred));
  return pred != 0;
}

template <int L, bool HTC>
__global__ void __launch_bounds__(QT_T, 1) qhead_tc_kernel(md_graphs_t G, md_state_t S, md_weights_t W, md_work_t K,
                                                          int* counter) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  QTSmem<L, HTC>& sm = *reinterpret_cast<QTSmem<L, HTC>*>(smem_raw);  // (no pointer arithmetic: keep the shared state space)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
  constexpr uint32_t NCOL = 64 * L;                    // accumulator columns per stage
  constexpr uint32_t TCOLS = (2 * NCOL <= 256 && !HTC) ? 256 : 512;
  constexpr uint32_t HCOL = 2 * NCOL;                  // HTC: 4 x 32 accumulator columns of relu's argument e_p @ h1
  const int tpg = G.stride / TILE_M;
  const int n_tiles = G.n_graphs * tpg;
  // Compact tiles: a tile is 128 consecutive rows of the concatenation of all graphs' active lists, whatever graphs they
  // belong to (every per-graph quantity of the epilogue is looked up per row).  Towards the end of a rollout, when most
  // graphs are down to a few dozen active nodes, this is a quarter of the (graph, 128-row) tiles -- and the CTAs beyond
  // the number of tiles leave before the set-up below (operand images of the weights, tensor memory).
  const bool compact = G.n_graphs <= QT_PMAX;
  int n_work = n_tiles;
  if (compact) {
    const int B = G.n_graphs;
    int carry = 0;
    for (int base = 0; base < B; base += QT_T) {
      const int gq = base + tid;
      const int v = gq < B ? max(S.n_act[gq], 0) : 0;
      int x = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
      }
      if (lane == 31) sm.wtot[warp] = x;
      __syncthreads();
      int woff = 0, tot = 0;
#pragma unroll
      for (int w2 = 0; w2 < QT_T / 32; ++w2) {
        const int tw = sm.wtot[w2];
        woff += w2 < warp ? tw : 0;
        tot += tw;
      }
      if (gq < B) sm.pfx[gq] = carry + woff + x - v;
      carry += tot;
      __syncthreads();
    }
    if (tid == 0) sm.pfx[B] = carry;
    n_work = (carry + TILE_M - 1) / TILE_M;
    if ((int)blockIdx.x >= n_work) {  // nothing for this CTA (tiles are handed out dynamically: n_work CTAs suffice)
      if (tid == 0) {
        __threadfence();
        if (atomicAdd(counter + 1, 1) == (int)gridDim.x - 1) {
          counter[0] = 0;
          counter[1] = 0;
        }
      }
      return;
    }
  }

  if (tid == 0) {
    mbar_init(&sm.a_full, 2);
    mbar_init(&sm.a_empty, 1);
    mbar_init(&sm.d_full[0], 1);
    mbar_init(&sm.d_full[1], 1);
    mbar_init(&sm.d_empty[0], 8);
    mbar_init(&sm.d_empty[1], 8);
    mbar_init(&sm.h_full, 1);
    mbar_fence_init();
  }
  if (warp == 10) tmem_alloc(&sm.tmem, TCOLS);
  // ---- weights: trans^T operand images, h1, bias, lw, h2 ------------------------------------------------------
  {
    float t[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) {
      const int i = tid + QT_T * j;
      t[j] = i < D * D ? W.trans[i] : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 12; ++j) {
      const int i = tid + QT_T * j;
      if (i < D * D) {
        const int k = i >> 6, n = i & 63, kb = k >> 5, kk = k & 31;
        const uint32_t off = sw128_off(n, kk >> 2) + 4 * (kk & 3);
        *reinterpret_cast<float*>(sm.Bt + kb * 8192 + off) = tf32_hi(t[j]);
        *reinterpret_cast<float*>(sm.Bt + 16384 + kb * 8192 + off) = t[j] - tf32_hi(t[j]);
        *reinterpret_cast<__half*>(sm.Bt + 32768 + sw128_off(n, k >> 3) + 2 * (k & 7)) = __float2half_rn(t[j]);
      }
    }
    if constexpr (HTC) {
      for (int i = tid; i < D * MD_HIDDEN; i += QT_T) {  // h1[c][j] -> B operand row j (hidden unit), K index c
        const int c = i / MD_HIDDEN, j = i - c * MD_HIDDEN, kb = c >> 5, kk = c & 31;
        const float x = W.h1[i];
        const uint32_t off = kb * 4096 + sw128_off(j, kk >> 2) + 4 * (kk & 3);
        *reinterpret_cast<float*>(sm.Bh + off) = tf32_rn(x);
        *reinterpret_cast<float*>(sm.Bh + 8192 + off) = tf32_rn(x - tf32_rn(x));
        *reinterpret_cast<__half*>(sm.Bh + 16384 + sw128_off(j, c >> 3) + 2 * (c & 7)) = __float2half_rn(x);
      }
    } else {
      for (int i = tid; i < D * MD_HIDDEN / 4; i += QT_T)
        reinterpret_cast<float4*>(sm.h1)[i] = reinterpret_cast<const float4*>(W.h1)[i];
    }
    if (tid < D) {
      sm.bias[tid] = W.bias[tid];
      sm.lw[tid] = W.logis_w[tid];
    }
    if (tid < MD_HIDDEN) sm.h2[tid] = W.h2[tid];
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = sm.tmem;
  const int hsel = MD_BP_ITERS & 1;

  if (warp == 8 || warp == 9) {
    // ================================ loaders ================================
    const int lw_ = warp - 8, ltid = tid - QT_EPI;
    for (uint32_t i = 0;; ++i) {
      const uint32_t slot = i & 1, use = i >> 1;
      if (ltid == 0) {  // next live tile
        int t;
        if (compact) {
          t = atomicAdd(counter, 1);
          if (t >= n_work) t = -1;
        } else {
          for (;;) {
            t = atomicAdd(counter, 1);
            if (t >= n_tiles) {
              t = -1;
              break;
            }
            const int g = t / tpg;
            if ((t - g * tpg) * TILE_M < S.n_act[g]) break;
          }
        }
        sm.next_tile = t;
      }
      asm volatile("bar.sync 1, 64;" ::: "memory");
      const int tile = sm.next_tile;
      // Row ids / descriptor of tile i live in slot i & 3 and need no wait: the loaders got here by storing the last layer of
      // tile i - 1, for which the MMAs of that tile's first layer had to be issued, which waited for the accumulator stage of
      // tile i - 3 to be released -- and the epilogue reads its row id and descriptor before it releases a stage.  So the
      // gather of tile i can start while the epilogues of tiles i - 2 and i - 1 are still running (with two slots it had to
      // wait for the END of epilogue i - 2 once T stayed in tensor memory: 18 % of the epilogue's time was spent waiting for
      // the next tile).
      const uint32_t ms = i & 3;
      (void)slot, (void)use;
      const int g = tile >= 0 && !compact ? tile / tpg : 0, tt = tile >= 0 && !compact ? tile - g * tpg : 0;
      if (tile >= 0 && compact) {
        const int B = G.n_graphs, total = sm.pfx[B];
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int j = tile * TILE_M + ltid + 64 * hh;
          int v = -1;
          if (j < total) {
            int lo = 0, hi = B;  // pfx[lo] <= j < pfx[hi]
            while (hi - lo > 1) {
              const int mid = (lo + hi) >> 1;
              if (sm.pfx[mid] <= j) lo = mid; else hi = mid;
            }
            v = S.act_list[lo * G.stride + (j - sm.pfx[lo])];
          }
          sm.vrow[ms][ltid + 64 * hh] = v;
        }
      } else if (tile >= 0) {
        const int rows = min(TILE_M, S.n_act[g] - tt * TILE_M);
        sm.vrow[ms][ltid] = ltid < rows ? S.act_list[g * G.stride + tt * TILE_M + ltid] : -1;
        sm.vrow[ms][ltid + 64] = ltid + 64 < rows ? S.act_list[g * G.stride + tt * TILE_M + ltid + 64] : -1;
      }
      if (ltid == 0) sm.desc[ms][0] = tile, sm.desc[ms][1] = g;
      asm volatile("bar.sync 1, 64;" ::: "memory");  // vrow visible to all loaders; next_tile may be overwritten
      if (tile < 0) {
        if (lane == 0) mbar_arrive(&sm.a_full);
        break;
      }
      const int c = lane & 15, half = lane >> 4;
#pragma unroll 1
      for (int l = 0; l < L; ++l) {
        const uint32_t ld = i * L + l;                       // operand-tile hand-off number
        const float* __restrict__ H = K.h[hsel][l];
        // all 32 row fetches of this lane (the warp's 64 rows) are in flight BEFORE the wait for the operand tiles: the load
        // latency overlaps the MMAs of the previous layer instead of following them
        float4 val[2][16];
#pragma unroll
        for (int hb = 0; hb < 2; ++hb) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int r = 64 * lw_ + 32 * hb + 2 * j + half;
            const int v = sm.vrow[ms][r];
            val[hb][j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (v >= 0) val[hb][j] = __ldg(reinterpret_cast<const float4*>(H + (size_t)v * D) + c);
          }
        }
        if (l + 1 < L) {  // next layer's rows of this tile: DRAM -> L2 now (two 128-byte lines per row, one row pair per lane)
          const float* __restrict__ Hn = K.h[hsel][l + 1];
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int v = sm.vrow[ms][64 * lw_ + 2 * lane + j];
            if (v >= 0) {
              asm volatile("prefetch.global.L2 [%0];" ::"l"(Hn + (size_t)v * D));
              asm volatile("prefetch.global.L2 [%0];" ::"l"(Hn + (size_t)v * D + 32));
            }
          }
        }
        if (ld >= 1) mbar_wait(&sm.a_empty, (ld - 1) & 1);   // the MMAs of the previous layer / tile have read the tiles
#pragma unroll
        for (int hb = 0; hb < 2; ++hb) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const uint32_t r = 64 * lw_ + 32 * hb + 2 * j + half;
            const float4 x = val[hb][j];
            *reinterpret_cast<float4*>(sm.A[c >> 3] + r * 128u + (((c & 7) ^ (r & 7)) << 4)) = x;
            const __half2 p0 = __floats2half2_rn(x.x - tf32_hi(x.x), x.y - tf32_hi(x.y));
            const __half2 p1 = __floats2half2_rn(x.z - tf32_hi(x.z), x.w - tf32_hi(x.w));
            *reinterpret_cast<uint2*>(sm.ALO + r * 128u + (((c >> 1) ^ (r & 7)) << 4) + 8 * (c & 1)) =
                make_uint2(*reinterpret_cast<const uint32_t*>(&p0), *reinterpret_cast<const uint32_t*>(&p1));
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.a_full);
      }
    }
  } else if (warp == 10) {
    // ================================ MMA issuer ================================
    const uint32_t idt = idesc(FMT_TF32, 128, 64), idb = idesc(FMT_F16, 128, 64);
    const uint32_t bt = smem_u32(sm.Bt);
    const uint64_t bh0 = desc_k_sw128(bt, 0), bh1 = desc_k_sw128(bt + 8192, 0);
    const uint64_t bl0 = desc_k_sw128(bt + 16384, 0), bl1 = desc_k_sw128(bt + 24576, 0);
    const uint64_t bbf = desc_k_sw128(bt + 32768, 0);
    const uint64_t a0 = desc_k_sw128(smem_u32(sm.A[0]), 0), a1 = desc_k_sw128(smem_u32(sm.A[1]), 0);
    const uint64_t alo = desc_k_sw128(smem_u32(sm.ALO), 0);
    for (uint32_t i = 0;; ++i) {
      const uint32_t slot = i & 1, use = i >> 1;
      mbar_wait(&sm.a_full, (i * L) & 1);  // first layer of the tile (or the end marker) is in place
      const int tile = sm.desc[i & 3][0];
      if (tile < 0) {
        if (lane == 0) mbar_arrive(&sm.d_full[slot]);
        break;
      }
      if (use >= 1) mbar_wait(&sm.d_empty[slot], (use - 1) & 1);  // accumulator stage drained
#pragma unroll 1
      for (int l = 0; l < L; ++l) {
        if (l > 0) mbar_wait(&sm.a_full, (i * L + l) & 1);
        tc_fence_after();
        if (qt_elect()) {
          const uint32_t d = tbase + slot * NCOL + 64 * l;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t ko = 2 * k;
            mma_tf32(d, a0 + ko, bh0 + ko, idt, k != 0);
            mma_tf32(d, a0 + ko, bl0 + ko, idt, 1);
            mma_tf32(d, a1 + ko, bh1 + ko, idt, 1);
            mma_tf32(d, a1 + ko, bl1 + ko, idt, 1);
            mma_f16(d, alo + ko, bbf + ko, idb, 1);
          }
          mma_commit(&sm.a_empty);
          if (l == L - 1) mma_commit(&sm.d_full[slot]);
        }
        __syncwarp();
      }
    }
  } else {
    // ================================ epilogue: two threads per row ================================
    // Warps w and w + 4 share TMEM lane quadrant w & 3.  The row's tanh outputs go to a column of shared memory
    // (Ts[l][c][row]: conflict-free) instead of registers: with everything unrolled the straight-line code no longer fit the
    // instruction cache (ncu: 44 % of the stall samples were instruction fetch), and rolled loops need an indexable T.
    // Half h of a row pair does the tanh of columns [32h, 32h+32) and the hidden units [16h, 16h+16) of relu(e @ h1);
    // the small reductions in between (gates, norms) are recomputed by both from the shared T.
    const float lb = W.logis_b[0];
    const int quad = warp & 3, half = warp >> 2;
    const int r = 32 * quad + lane;
    const uint32_t h1_a = smem_u32(sm.h1);
    uint32_t hph = 0;  // HTC: uses of h_full so far
    for (uint32_t i = 0;; ++i) {
      const uint32_t slot = i & 1, use = i >> 1;
      mbar_wait(&sm.d_full[slot], use & 1);
      const int tile = sm.desc[i & 3][0];
      if (tile < 0) break;
      tc_fence_after();
      const int v = sm.vrow[i & 3][r];
      const int g = v >= 0 ? v / G.stride : -1;  // the row's graph (rows of several graphs share a compact tile)
      float part[L];
#pragma unroll
      for (int p = 0; p < L; ++p) part[p] = 0.f;
      float gwp[L], aqp[L];
      if constexpr (HTC) {
        // ---- T = tanh(acc + bias) written back over the accumulators; partial gate dot products from the registers ----
        constexpr int NP = L * (L + 1) / 2;
        const uint32_t trow = tbase + ((uint32_t)(32 * quad) << 16) + slot * NCOL + 32 * half;  // this thread's 32 columns of layer 0
        float dp[NP];
#pragma unroll
        for (int i = 0; i < NP; ++i) dp[i] = 0.f;
#pragma unroll 1
        for (int j = 0; j < 2; ++j) {
          float t[L][16];
#pragma unroll
          for (int l = 0; l < L; ++l) tmem_ld16(trow + 64 * l + 16 * j, t[l]);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float b = sm.bias[32 * half + 16 * j + i];
#pragma unroll
            for (int l = 0; l < L; ++l) t[l][i] = tanh_fast(t[l][i] + b);
          }
#pragma unroll
          for (int l = 0; l < L; ++l) tmem_st16(trow + 64 * l + 16 * j, t[l]);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float w1 = sm.lw[32 * half + 16 * j + i];
            int idx = 0;
#pragma unroll
            for (int p = 0; p < L; ++p)
#pragma unroll
              for (int k = 0; k <= p; ++k, ++idx) dp[idx] = fmaf(t[k][i] * t[p][i], w1, dp[idx]);
          }
        }
        tmem_st_wait();
#pragma unroll
        for (int i = 0; i < NP; ++i) sm.dpart[half][i][r] = dp[i];
        asm volatile("bar.sync 2, 256;" ::: "memory");  // both halves of the gate dot products are in shared memory
        float am[L][L];
        {
          float dot[L][L];
          int idx = 0;
#pragma unroll
          for (int p = 0; p < L; ++p)
#pragma unroll
            for (int k = 0; k <= p; ++k, ++idx) dot[p][k] = dot[k][p] = sm.dpart[0][idx][r] + sm.dpart[1][idx][r];
#pragma unroll
          for (int p = 0; p < L; ++p) {
            float mx = -1e30f, den = 0.f;
#pragma unroll
            for (int k = 0; k < L; ++k) {
              am[p][k] = 1.f / (1.f + expf(-(dot[p][k] + lb)));
              mx = fmaxf(mx, am[p][k]);
            }
#pragma unroll
            for (int k = 0; k < L; ++k) {
              am[p][k] = expf(am[p][k] - mx);
              den += am[p][k];
            }
#pragma unroll
            for (int k = 0; k < L; ++k) am[p][k] = am[p][k] / den;
          }
        }
        // X_p = T_p + sum_{k != p} a_k T_k as sum_k cf[k] T_k (cf[p] = 1) over this thread's 32 columns, re-read from tensor
        // memory: the UNSCALED rows go to the operand tiles and the scale 1/|X_p| * s_p is applied to the 32 products afterwards
        // (the norm needs both threads' halves, exchanged through shared memory across the barrier that is there anyway).
        // Two operand buffers: X_{p+1} is written while the tensor core works on X_p (the wait for those ~20 small MMAs was
        // 14 % of the epilogue's time when it was exposed once per layer).
        float sgv[L];
        auto write_x = [&](int p) {
          const int buf = p & 1;
          float cf[L];
#pragma unroll
          for (int k = 0; k < L; ++k) {
            cf[k] = 0.f;
#pragma unroll
            for (int pp = 0; pp < L; ++pp)
              if (pp == p) cf[k] = (k == pp) ? 1.f : am[pp][k];
          }
          float ssh = 0.f;
          sgv[p] = 0.f, gwp[p] = 0.f, aqp[p] = 0.f;
          if (v >= 0) {
            sgv[p] = K.gconst[(size_t)g * MD_GCONST + p];
            gwp[p] = K.gconst[(size_t)g * MD_GCONST + 4 + p];
            aqp[p] = K.gconst[(size_t)g * MD_GCONST + 8 + p];
          }
          unsigned char* erow = sm.E[buf][half] + (uint32_t)r * 128u;
          unsigned char* lrow = sm.ELO[buf] + (uint32_t)r * 128u;
#pragma unroll 1
          for (int j = 0; j < 2; ++j) {
            float t[L][16];
#pragma unroll
            for (int l = 0; l < L; ++l) tmem_ld16(trow + 64 * l + 16 * j, t[l]);
            tmem_ld_wait();
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {  // 4 columns = one 16-byte chunk of the K block
              const int ch = 4 * j + q4;
              float e[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                float x = cf[0] * t[0][4 * q4 + i];
#pragma unroll
                for (int k = 1; k < L; ++k) x = fmaf(cf[k], t[k][4 * q4 + i], x);
                ssh = fmaf(x, x, ssh);
                e[i] = x;
              }
              const float4 eh = make_float4(tf32_rn(e[0]), tf32_rn(e[1]), tf32_rn(e[2]), tf32_rn(e[3]));
              *reinterpret_cast<float4*>(erow + ((ch ^ (r & 7)) << 4)) = eh;
              // remainders scaled by 2^12 (exact): unscaled they reach fp16's subnormal range with a handful of bits left
              // (measured with the scaled rows e_p, |e| ~ 1e-3: 1.3e-6 absolute Q error); their product has its own accumulator
              const __half2 p0 = __floats2half2_rn((e[0] - eh.x) * 4096.f, (e[1] - eh.y) * 4096.f);
              const __half2 p1 = __floats2half2_rn((e[2] - eh.z) * 4096.f, (e[3] - eh.w) * 4096.f);
              // fp16 row: column c at byte 2 c -> chunk c >> 3 = 4 half + (ch >> 1), 8 bytes at (ch & 1) * 8
              *reinterpret_cast<uint2*>(lrow + (((4 * half + (ch >> 1)) ^ (r & 7)) << 4) + 8 * (ch & 1)) =
                  make_uint2(*reinterpret_cast<const uint32_t*>(&p0), *reinterpret_cast<const uint32_t*>(&p1));
            }
          }
          sm.ssq[buf][half][r] = ssh;
        };
        write_x(0);
#pragma unroll 1
        for (int p = 0; p < L; ++p) {
          fence_proxy_async_smem();
          tc_fence_before();
          asm volatile("bar.sync 3, 256;" ::: "memory");  // X_p complete; the previous layer's accumulator reads are done
          if (warp == 0) {
            tc_fence_after();
            if (qt_elect()) {
              const uint32_t idt = idesc(FMT_TF32, 128, 32), idb = idesc(FMT_F16, 128, 32);
              const uint32_t bh = smem_u32(sm.Bh);
              const uint32_t d = tbase + HCOL;
#pragma unroll
              for (int kb = 0; kb < 2; ++kb) {
                const uint64_t ae = desc_k_sw128(smem_u32(sm.E[p & 1][kb]), 0);
                const uint64_t bhi = desc_k_sw128(bh + kb * 4096, 0), blo = desc_k_sw128(bh + 8192 + kb * 4096, 0);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  // four accumulators: the tensor core truncates when it adds into an accumulator, a bias that grows with
                  // the number of steps (16 steps into one accumulator: 1.2e-6 absolute Q error where |q| ~ 1e-4; the
                  // FP32 pipes: 4.5e-7).  hi x hi gets one accumulator per K block (4 steps each), the two small
                  // products their own; the epilogue adds the four in FP32.
                  mma_tf32(d + 32 * kb, ae + 2 * k, bhi + 2 * k, idt, k != 0);
                  mma_tf32(d + 64, ae + 2 * k, blo + 2 * k, idt, (kb | k) != 0);
                }
              }
              const uint64_t al = desc_k_sw128(smem_u32(sm.ELO[p & 1]), 0), bf = desc_k_sw128(bh + 16384, 0);
#pragma unroll
              for (int k = 0; k < 4; ++k) mma_f16(d + 96, al + 2 * k, bf + 2 * k, idb, k != 0);
              mma_commit(&sm.h_full);
            }
            __syncwarp();
          }
          // the other buffer is free: its MMAs (layer p - 1) completed before that layer's products were read, and every
          // thread read its ssq[.] entries before the barrier above
          if (p + 1 < L) write_x(p + 1);
          mbar_wait(&sm.h_full, hph & 1);
          ++hph;
          tc_fence_after();
          float x[16], xb[16], xm[16], xl[16];
          const uint32_t hta = tbase + ((uint32_t)(32 * quad) << 16) + HCOL + 16 * half;
          tmem_ld16(hta, x);
          tmem_ld16(hta + 32, xb);
          tmem_ld16(hta + 64, xm);
          tmem_ld16(hta + 96, xl);
          tmem_ld_wait();
          const float ss = sm.ssq[p & 1][0][r] + sm.ssq[p & 1][1][r];
          const float scale = __frcp_rn(fmaxf(sqrtf(ss), 1e-12f)) * sgv[p];  // e_p = X_p / |X_p| * s_p
          float hq = 0.f;
#pragma unroll
          for (int j = 0; j < MD_HIDDEN / 2; ++j)
            hq = fmaf(fmaxf(((x[j] + xb[j]) + fmaf(xl[j], 1.f / 4096.f, xm[j])) * scale, 0.f), sm.h2[16 * half + j], hq);
          part[p] = hq;
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.d_empty[slot]);  // T (and the accumulators under it) consumed: release the stage
      } else {
#pragma unroll 1
      for (int lc = 0; lc < L * 2; ++lc) {
        const int l = lc >> 1, c0 = 32 * half + (lc & 1) * 16;
        float x[16];
        tmem_ld16(tbase + ((uint32_t)(32 * quad) << 16) + slot * NCOL + 64 * l + c0, x);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) sm.Ts[l][c0 + j][r] = tanhf(x[j] + sm.bias[c0 + j]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.d_empty[slot]);  // accumulators consumed: release the stage early
      asm volatile("bar.sync 2, 256;" ::: "memory");  // both column halves of T are in shared memory
      if (v >= 0) {
        float dot[L][L];
#pragma unroll
        for (int p = 0; p < L; ++p)
#pragma unroll
          for (int k = 0; k < L; ++k) dot[p][k] = 0.f;
#pragma unroll 4
        for (int c = 0; c < D; ++c) {
          float t[L];
#pragma unroll
          for (int l = 0; l < L; ++l) t[l] = sm.Ts[l][c][r];
          const float w1 = sm.lw[c];
#pragma unroll
          for (int p = 0; p < L; ++p)
#pragma unroll
            for (int k = 0; k <= p; ++k) dot[p][k] = fmaf(t[k] * t[p], w1, dot[p][k]);
        }
#pragma unroll
        for (int p = 0; p < L; ++p)
#pragma unroll
          for (int k = 0; k < p; ++k) dot[k][p] = dot[p][k];
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
#pragma unroll
          for (int k = 0; k < L; ++k) a[k] = a[k] / den;
          // X_p = T_p + sum_{k != p} a_k T_k ; |X_p| ; then e_p = X_p / |X_p| * s_p feeds relu(e_p @ h1) . h2
          float ss = 0.f;
#pragma unroll 4
          for (int c = 0; c < D; ++c) {
            float x = sm.Ts[p][c][r];
#pragma unroll
            for (int k = 0; k < L; ++k)
              if (k != p) x = fmaf(a[k], sm.Ts[k][c][r], x);
            ss = fmaf(x, x, ss);
          }
          const float sc = __frcp_rn(fmaxf(sqrtf(ss), 1e-12f));
          const float sg = K.gconst[(size_t)g * MD_GCONST + p];
          gwp[p] = K.gconst[(size_t)g * MD_GCONST + 4 + p];
          aqp[p] = K.gconst[(size_t)g * MD_GCONST + 8 + p];
          float hid[MD_HIDDEN / 2];
#pragma unroll
          for (int j = 0; j < MD_HIDDEN / 2; ++j) hid[j] = 0.f;
#pragma unroll 2
          for (int c = 0; c < D; ++c) {
            float x = sm.Ts[p][c][r];
#pragma unroll
            for (int k = 0; k < L; ++k)
              if (k != p) x = fmaf(a[k], sm.Ts[k][c][r], x);
            const float e = x * sc * sg;
#pragma unroll
            for (int j4 = 0; j4 < MD_HIDDEN / 8; ++j4) {
              const float4 h = qt_lds128(h1_a + 4 * (c * MD_HIDDEN + 16 * half + 4 * j4));
              hid[4 * j4] = fmaf(e, h.x, hid[4 * j4]);
              hid[4 * j4 + 1] = fmaf(e, h.y, hid[4 * j4 + 1]);
              hid[4 * j4 + 2] = fmaf(e, h.z, hid[4 * j4 + 2]);
              hid[4 * j4 + 3] = fmaf(e, h.w, hid[4 * j4 + 3]);
            }
          }
          float hq = 0.f;
#pragma unroll
          for (int j = 0; j < MD_HIDDEN / 2; ++j) hq = fmaf(fmaxf(hid[j], 0.f), sm.h2[16 * half + j], hq);
          part[p] = hq;
        }
      }
      }
      if (half == 1) {
#pragma unroll
        for (int p = 0; p < L; ++p) sm.hq[p][r] = part[p];
      }
      asm volatile("bar.sync 2, 256;" ::: "memory");  // partial hidden sums exchanged; every Ts read of this tile is done
      unsigned long long key = 0ull;
      if (half == 0 && v >= 0) {
        float q = 0.f;
#pragma unroll
        for (int p = 0; p < L; ++p) q = fmaf(gwp[p], (part[p] + sm.hq[p][r]) + aqp[p], q);
        K.q[v] = q;
        const uint32_t local = (uint32_t)(v - g * G.stride);
        key = ((unsigned long long)orderable(q) << 32) | (0xffffffffu - local);
      }
      if (half == 0) {
        // arg-max per graph: a graph's rows are consecutive lanes, so a segmented suffix maximum leaves every graph's
        // best key in the first of its lanes
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const unsigned long long other = __shfl_down_sync(0xffffffffu, key, o);
          const int og = __shfl_down_sync(0xffffffffu, g, o);
          if (lane + o < 32 && og == g && other > key) key = other;
        }
        const int gprev = __shfl_up_sync(0xffffffffu, g, 1);
        if (g >= 0 && (lane == 0 || gprev != g) && key != 0ull) atomicMax(&S.best[g], key);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 10) tmem_dealloc(tbase, TCOLS);
  // the last CTA to leave re-arms the tile queue for the next launch (no memset node between launches)
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(counter + 1, 1) == (int)gridDim.x - 1) {
      counter[0] = 0;
      counter[1] = 0;
    }
  }
}

}  // namespace md

extern int g_md_debug_flags;

// *launched = 1 when the tensor-core kernel was enqueued, 0 when the caller must use the FP32-FMA kernel; returns an error code
int md_qhead_tc_launch(const md_graphs_t* g, const md_state_t* s, const md_weights_t* w, const md_work_t* k, cudaStream_t st,
                       int* launched) {
  *launched = 0;
  if ((g_md_debug_flags & 8) || !s->flags || (g->n_layers != 2 && g->n_layers != 3)) return 0;
  static md::PerDevice<int> nsm_pd;
  int& nsm = nsm_pd.get();
  if (nsm <= 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    if (nsm <= 0) nsm = 148;
  }
  const int n_tiles = g->n_graphs * (g->stride / MD_TILE_M);
  const int grid = nsm < n_tiles ? nsm : n_tiles;
  int* counter = s->flags + 4;  // [4] next tile, [5] CTAs finished; zero between launches (the kernel re-arms them)
  cudaError_t e = cudaSuccess;
  if (g->n_layers == 2 && !(g_md_debug_flags & 8192)) {  // (bit 13: hidden layer on the FP32 pipes, the A/B reference)
    static md::PerDevice<bool> configured_pd;
  bool& configured = configured_pd.get();
    const int smem = (int)sizeof(md::QTSmem<2, true>);
    if (!configured) {
      e = cudaFuncSetAttribute(md::qhead_tc_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      if (e != cudaSuccess) return (int)e;
      configured = true;
    }
    md::qhead_tc_kernel<2, true><<<grid, md::QT_T, smem, st>>>(*g, *s, *w, *k, counter);
  } else if (g->n_layers == 2) {
    static md::PerDevice<bool> configured_pd;
  bool& configured = configured_pd.get();
    const int smem = (int)sizeof(md::QTSmem<2, false>);
    if (!configured) {
      e = cudaFuncSetAttribute(md::qhead_tc_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      if (e != cudaSuccess) return (int)e;
      configured = true;
    }
    md::qhead_tc_kernel<2, false><<<grid, md::QT_T, smem, st>>>(*g, *s, *w, *k, counter);
  } else if (!(g_md_debug_flags & 8192)) {
    static md::PerDevice<bool> configured_pd;
    bool& configured = configured_pd.get();
    const int smem = (int)sizeof(md::QTSmem<3, true>);
    if (!configured) {
      e = cudaFuncSetAttribute(md::qhead_tc_kernel<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      if (e != cudaSuccess) return (int)e;
      configured = true;
    }
    md::qhead_tc_kernel<3, true><<<grid, md::QT_T, smem, st>>>(*g, *s, *w, *k, counter);
  } else {
    static md::PerDevice<bool> configured_pd;
  bool& configured = configured_pd.get();
    const int smem = (int)sizeof(md::QTSmem<3, false>);
    if (!configured) {
      e = cudaFuncSetAttribute(md::qhead_tc_kernel<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      if (e != cudaSuccess) return (int)e;
      configured = true;
    }
    md::qhead_tc_kernel<3, false><<<grid, md::QT_T, smem, st>>>(*g, *s, *w, *k, counter);
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  *launched = 1;
  return 0;
}
