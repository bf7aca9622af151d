// md_agg: neighbour aggregation for graphs whose embeddings do not fit on chip (the N = 1M configuration):
//   P[v, :] = sum over the alive neighbours u of v (CSR order) of H[u, :]      -- torch_sparse.spmm(n2nsum_param, H)
// (/root/reference/code/MultiDismantler_unit_cost/MultiDismantler_net_graphsage.py:290), written into the rows of the
// OTHER ping-pong embedding buffer, where the dense-transform kernel (md_embed_iter_tc.cu, PRE mode) picks it up, tile by
// tile, and overwrites it with the new embeddings.
//
// Why a kernel of its own: with H_l = 256 MB the gather is a stream of random 256-byte row reads from HBM.  Measured on
// B200 (scripts/microbench/gather_bw.cu): such reads reach 7.4-9.1 TB/s of gathered rows (L2 catches part of a 256 MB
// table) once >= 64 KB per SM are in flight; the fused iteration kernel, which shares the SM with the transform (16
// gather warps, one row pair at a time, shared memory full of operand tiles), sustained 2.3 TB/s.
//
// agg_ring_kernel (default): 8 warps per SM, every warp streams the neighbour rows of its 32-row chunks through a private
// shared-memory ring with cp.async (16 bytes per lane, one 256-byte row per half-warp, no registers held while the bytes
// are in flight): 3 slots of 2 x 16 rows, two slots (16 KB per warp, 128 KB per SM) always in flight, independent of the
// row lengths -- the entries of a half-warp's 16 destination rows form one flat stream, summed from shared memory in CSR
// order with a flush at every row boundary.  Descriptors (row pointers, node ids, neighbour ids) of the next chunk are
// fetched while the ring is full.
// agg_rows_kernel: the register-load version (32 warps, 8 loads per lane), kept as the A/B reference (md_debug_set_flags
// bit 9): 1.69 ms per launch at C4 against the ring's figure in DESIGN 6.
#include "md_common.cuh"

namespace md {

constexpr int AG_T = 1024;
constexpr int AG_W = AG_T / 32;
constexpr int AG_IDS = 352;  // staged neighbour ids of one 32-row chunk (mean 256 at degree 8); larger chunks read L2

__global__ void __launch_bounds__(AG_T, 1) agg_rows_kernel(md_graphs_t G, md_state_t S, md_work_t K, int iter) {
  __shared__ int ids[AG_W][AG_IDS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, half = lane >> 4, l16 = lane & 15;
  const int St = G.stride, B = G.n_graphs, L = G.n_layers;
  const int cpg = St / 32;  // 32-row chunks per graph (the stride is a multiple of 128)
  const long long per_layer = (long long)B * cpg, n_chunks = per_layer * L;
  for (long long c = (long long)blockIdx.x * AG_W + warp; c < n_chunks; c += (long long)gridDim.x * AG_W) {
    const int l = (int)(c / per_layer);
    const int rem = (int)(c - (long long)l * per_layer);
    const int g = rem / cpg, p0 = (rem - g * cpg) * 32;
    const int n = S.n_act[g];
    if (p0 >= n) continue;  // warp-uniform
    const int rows = min(32, n - p0);
    const int32_t* __restrict__ aptr = S.aptr[l] + (size_t)g * St + p0;
    const int32_t* __restrict__ acol = S.acol[l];
    const int ap = aptr[min(lane, rows)], ap_hi = aptr[min(lane + 1, rows)];
    const int v = lane < rows ? S.act_list[(size_t)g * St + p0 + lane] : -1;
    const int e_begin = __shfl_sync(0xffffffffu, ap, 0);
    const int total = __shfl_sync(0xffffffffu, ap_hi, 31) - e_begin;
    const bool staged = total <= AG_IDS;
    if (staged)
      for (int k = lane; k < total; k += 32) ids[warp][k] = acol[e_begin + k];
    __syncwarp();
    const float4* __restrict__ Hin = reinterpret_cast<const float4*>(K.h[iter & 1][l]);
    float4* __restrict__ Pout = reinterpret_cast<float4*>(K.h[(iter + 1) & 1][l]);
#pragma unroll 1
    for (int i = 0; 2 * i < rows; ++i) {
      const int r = 2 * i + half;
      const int e0 = __shfl_sync(0xffffffffu, ap, r), e1 = __shfl_sync(0xffffffffu, ap_hi, r);
      const int vv = __shfl_sync(0xffffffffu, v, r);
      const int nr = r < rows ? e1 - e0 : 0;
      const int nmax = max(nr, __shfl_xor_sync(0xffffffffu, nr, 16));
      const int* my = &ids[warp][e0 - e_begin];
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int k0 = 0; k0 < nmax; k0 += 8) {
        float4 val[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          val[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (k0 + j < nr) {
            const int u = staged ? my[k0 + j] : acol[e0 + k0 + j];
            val[j] = __ldg(Hin + (size_t)u * (D / 4) + l16);
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {  // CSR order: the f32 sum is deterministic and equals a sequential scatter-add
          if (k0 + j < nr) {
            acc.x += val[j].x;
            acc.y += val[j].y;
            acc.z += val[j].z;
            acc.w += val[j].w;
          }
        }
      }
      if (vv >= 0) Pout[(size_t)vv * (D / 4) + l16] = acc;
    }
    __syncwarp();
  }
}


// ---- cp.async ring version ---------------------------------------------------------------------------------------------
constexpr int RG_W = 8;
constexpr int RG_T = RG_W * 32;
constexpr int RG_R = 3;      // ring slots per warp
constexpr int RG_IDS = 304;  // staged neighbour entries of one chunk (mean 256 at degree 8, sd 16)
constexpr int RG_DB = 3;     // descriptor buffers per warp (the issue cursor runs at most two chunks ahead)
constexpr int RG_IDBITS = 27;  // entry = neighbour id | destination row (0..31) << 27
constexpr uint32_t RG_SLOT_B = 8192;

struct RingDesc {
  int ent[RG_IDS];  // staged entries (see RG_IDBITS), CSR order
  int vrow[32];     // node (global id) of the row, -1 = none
  int total, layer, pad[2];
};
struct RingSmem {
  unsigned char ring[RG_W][RG_R][RG_SLOT_B];  // [warp][slot][32 entries x 256 bytes]
  RingDesc desc[RG_W][RG_DB];
  int fifo[RG_W][4];  // per in-flight slot: descriptor buffer | step << 8, -1 = nothing issued
};
static_assert(sizeof(RingSmem) <= 232448, "RingSmem exceeds the shared memory of one CTA");

__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ uint64_t lds_b64(uint32_t a) {
  uint64_t v;
  asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void add_f32x2(uint64_t& a, uint64_t b) { asm("add.rn.f32x2 %0, %0, %1;" : "+l"(a) : "l"(b)); }

// One WARP per stream: the entries of a chunk's 32 destination rows form one flat list, cut into steps of 32 entries; a
// step is 32 rows of 256 bytes, every lane copying / summing 8 bytes of each (all control flow is warp-uniform).
__global__ void __launch_bounds__(RG_T, 1) agg_ring_kernel(md_graphs_t G, md_state_t S, md_work_t K, int iter) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  RingSmem& sm = *reinterpret_cast<RingSmem*>(smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int St = G.stride, B = G.n_graphs, L = G.n_layers;
  const int cpg = St / 32;  // 32-row chunks per graph (the stride is a multiple of 128)
  const long long per_layer = (long long)B * cpg, n_chunks = per_layer * L;
  const long long cstride = (long long)gridDim.x * RG_W;
  long long c_next = (long long)blockIdx.x * RG_W + warp;  // next chunk index to look at
  // this lane's 8-byte column pair of the input / output embedding rows, per layer (no dynamic indexing of the params)
  auto hin_of = [&](int l) -> const float2* {
    const float* p = K.h[iter & 1][0];
    if (l == 1) p = K.h[iter & 1][1];
    if (l == 2) p = K.h[iter & 1][2];
    if (l == 3) p = K.h[iter & 1][3];
    return reinterpret_cast<const float2*>(p) + lane;
  };
  auto pout_of = [&](int l) -> uint64_t* {
    float* p = K.h[(iter + 1) & 1][0];
    if (l == 1) p = K.h[(iter + 1) & 1][1];
    if (l == 2) p = K.h[(iter + 1) & 1][2];
    if (l == 3) p = K.h[(iter + 1) & 1][3];
    return reinterpret_cast<uint64_t*>(p) + lane;
  };
  auto acol_of = [&](int l) -> const int32_t* {
    const int32_t* p = S.acol[0];
    if (l == 1) p = S.acol[1];
    if (l == 2) p = S.acol[2];
    if (l == 3) p = S.acol[3];
    return p;
  };
  auto aptr_of = [&](int l) -> const int32_t* {
    const int32_t* p = S.aptr[0];
    if (l == 1) p = S.aptr[1];
    if (l == 2) p = S.aptr[2];
    if (l == 3) p = S.aptr[3];
    return p;
  };

  // Fetch the descriptors of the warp's next chunk into buffer b; false when the sequence is exhausted.  Rows without
  // entries get their zeros here; chunks whose entries do not fit the staging area (hub rows) are summed right here with
  // register loads and skipped.
  auto load_desc = [&](int b) -> bool {
    for (; c_next < n_chunks; c_next += cstride) {
      const int l = (int)(c_next / per_layer);
      const int rem = (int)(c_next - (long long)l * per_layer);
      const int g = rem / cpg, p0 = (rem - g * cpg) * 32;
      const int n = S.n_act[g];
      if (p0 >= n) continue;
      const int rows = min(32, n - p0);
      const int32_t* __restrict__ aptr = aptr_of(l) + (size_t)g * St + p0;
      const int ap = aptr[min(lane, rows)], ap_hi = aptr[min(lane + 1, rows)];
      const int v = lane < rows ? S.act_list[(size_t)g * St + p0 + lane] : -1;
      const int e_begin = __shfl_sync(0xffffffffu, ap, 0);
      const int total = __shfl_sync(0xffffffffu, ap_hi, 31) - e_begin;
      const int32_t* __restrict__ acol = acol_of(l) + e_begin;
      uint64_t* __restrict__ Pout = pout_of(l);
      if (total > RG_IDS) {
        const float2* __restrict__ Hin = hin_of(l);
        for (int r = 0; r < rows; ++r) {
          const int e0 = __shfl_sync(0xffffffffu, ap, r) - e_begin, e1 = __shfl_sync(0xffffffffu, ap_hi, r) - e_begin;
          const int vv = __shfl_sync(0xffffffffu, v, r);
          float2 acc = make_float2(0.f, 0.f);
          for (int k0 = e0; k0 < e1; k0 += 8) {
            float2 val[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              val[j] = make_float2(0.f, 0.f);
              if (k0 + j < e1) val[j] = __ldg(Hin + (size_t)acol[k0 + j] * (D / 2));
            }
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (k0 + j < e1) acc.x += val[j].x, acc.y += val[j].y;
          }
          reinterpret_cast<float2*>(Pout)[(size_t)vv * (D / 2)] = acc;
        }
        continue;
      }
      RingDesc& d = sm.desc[warp][b];
      d.vrow[lane] = v;
      if (lane == 0) d.total = total, d.layer = l;
      for (int k = lane; k < total; k += 32) d.ent[k] = acol[k];
      if (v >= 0 && ap_hi == ap) Pout[(size_t)v * (D / 2)] = 0ull;  // (an active node has alive neighbours: defensive)
      __syncwarp();
      for (int e = ap - e_begin; e < ap_hi - e_begin; ++e) d.ent[e] |= lane << RG_IDBITS;  // lane = destination row
      __syncwarp();
      c_next += cstride;
      return true;
    }
    return false;
  };
  // number of 32-entry steps of the chunk in buffer b (>= 1 keeps the stream bookkeeping uniform)
  auto n_steps = [&](int b) -> int { return max(1, (sm.desc[warp][b].total + 31) >> 5); };
  const uint32_t ring_a = (uint32_t)__cvta_generic_to_shared(&sm.ring[warp][0][8 * lane]);
  // issue: 16 bytes per lane, L1 bypassed (.cg) -- half-warp h copies entry 2 j + h; the consumer reads the slot with its
  // own mapping (8 bytes of every entry per lane), so the two sides need not agree
  const int half = lane >> 4, l16 = lane & 15;
  const uint32_t ring_i = (uint32_t)__cvta_generic_to_shared(&sm.ring[warp][0][256 * half + 16 * l16]);
  auto issue = [&](int b, int si, int slot) {
    const RingDesc& d = sm.desc[warp][b];
    const int nv = d.total - si * 32;
    const int w = lane < nv ? d.ent[si * 32 + lane] : 0;
    const float4* __restrict__ Hin = reinterpret_cast<const float4*>(hin_of(d.layer) - lane) + l16;
    const uint32_t dst = ring_i + (uint32_t)slot * RG_SLOT_B;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int wj = __shfl_sync(0xffffffffu, w, 2 * j + half);
      if (2 * j + half < nv) cp_async16(dst + 512u * j, Hin + (size_t)(wj & ((1 << RG_IDBITS) - 1)) * (D / 4));
    }
  };

  // consumer state: destination row being summed (0..31 within the chunk), its accumulator (two packed floats per lane)
  int row = 0;
  uint64_t acc = 0ull;
  uint64_t* __restrict__ Pout = nullptr;
  auto begin_chunk = [&](int b) {
    const RingDesc& d = sm.desc[warp][b];
    row = d.total > 0 ? (int)((unsigned)d.ent[0] >> RG_IDBITS) : -1;
    acc = 0ull;
    Pout = pout_of(d.layer);
  };
  auto consume = [&](int b, int si, int slot, bool last) {
    const RingDesc& d = sm.desc[warp][b];
    const int nv = d.total - si * 32;
    const int w = lane < nv ? d.ent[si * 32 + lane] : 0;
    const int rj = (int)((unsigned)w >> RG_IDBITS);
    int rprev = __shfl_up_sync(0xffffffffu, rj, 1);
    if (lane == 0) rprev = row;
    const unsigned starts = __ballot_sync(0xffffffffu, lane < nv && rj != rprev);  // entries that open a new row
    const uint32_t src = ring_a + (uint32_t)slot * RG_SLOT_B;
    uint64_t x[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) x[j] = j < nv ? lds_b64(src + 256u * j) : 0ull;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if ((starts >> j) & 1u) {  // warp-uniform: the row that ends here leaves, the next one starts
        Pout[(size_t)d.vrow[row] * (D / 2)] = acc;
        acc = 0ull;
        row = __shfl_sync(0xffffffffu, rj, j);
      }
      add_f32x2(acc, x[j]);  // (+0 beyond the end of the stream)
    }
    if (last && row >= 0) Pout[(size_t)d.vrow[row] * (D / 2)] = acc;
  };

  // ---- the step stream: issue runs RG_R - 1 steps ahead of consume -----------------------------------------------
  // Chunk k uses descriptor buffer k % RG_DB.  The issue cursor fetches the next chunk's descriptors lazily, right before
  // its first step is issued: by then the consume cursor is at most one chunk behind, so the buffer being overwritten
  // (chunk k - 2's) is no longer read.
  int ib = 0, is = 0, in = 0;  // issue cursor: descriptor buffer, step, steps of that chunk
  int cs = 0, cn = 0;          // consume cursor: step, steps of its chunk
  if (!load_desc(0)) return;
  in = n_steps(0);
  cn = in;
  begin_chunk(0);
  bool issue_done = false;
  unsigned q = 0;  // stream positions issued so far (ring slot = position % RG_R)
  auto issue_next = [&]() {
    int tag = -1;
    if (!issue_done && is == in) {
      const int nb = (ib + 1) % RG_DB;
      if (load_desc(nb)) ib = nb, is = 0, in = n_steps(nb);
      else issue_done = true;
    }
    if (!issue_done) {
      issue(ib, is, (int)(q % RG_R));
      tag = ib | (is << 8);
      ++is;
    }
    cp_async_commit();
    if (lane == 0) sm.fifo[warp][q % RG_R] = tag;
    ++q;
  };
  for (int k = 0; k < RG_R - 1; ++k) issue_next();
  for (unsigned qc = 0;; ++qc) {
    issue_next();
    cp_async_wait<RG_R - 1>();
    __syncwarp();
    const int tag = sm.fifo[warp][qc % RG_R];
    if (tag < 0) break;  // nothing was issued for this position: the stream is exhausted
    const int b = tag & 0xff;
    consume(b, tag >> 8, (int)(qc % RG_R), cs + 1 == cn);
    if (++cs == cn) {  // next chunk (its first step is already in flight, so its descriptors are in place)
      const int nt = sm.fifo[warp][(qc + 1) % RG_R];
      if (nt >= 0) {
        cs = 0;
        cn = n_steps(nt & 0xff);
        begin_chunk(nt & 0xff);
      }
    }
    __syncwarp();
  }
  cp_async_wait<0>();
}


// ---- iterations 0 and 1, unit cost: no embedding gather at all ------------------------------------------------------------
// The first embedding of a node is a function of its alive degree only: relu([d/dmax, d/dmax] @ w_n2l), normalised
// (MultiDismantler_net_graphsage.py:248-287 with the unit-cost features of PrepareBatchGraph.py).  So the first neighbour
// sum needs no 256-byte row per neighbour: a table of the <= DG_CAP + 1 distinct rows of the (layer, graph) lives in shared
// memory (built with exactly embed_init_kernel's arithmetic) and the gather reads 4 bytes (the neighbour's alive degree,
// L2-resident) per entry.  The rows are still added one by one in CSR order.  md_embed_init is skipped altogether on this
// path; the transform kernel forms its own-row operand from the degrees as well (md_embed_iter_tc.cu).
//
// LEVEL 1 goes one step further.  relu([x, x] @ w_n2l) = x relu(w_n2l[0] + w_n2l[1]) for x > 0, so after the normalise
// EVERY node with an alive neighbour has the same first embedding c = relu(w0 + w1) / |relu(w0 + w1)| (to the last bit
// or two of the rounding of x), the first neighbour sum is deg(v) c, and the embedding after iteration 0 is again a function
// of the alive degree only:  H1(d) = normalise(relu(d (c @ W13) + c @ W23)).  Iteration 0 therefore needs no kernel at all,
// and iteration 1's neighbour sum is this kernel with the table H1(d) (one table for all layers and graphs) instead of a
// 256-byte gather per neighbour.  Deviation from adding the rows one by one: a few 1e-7 relative (parity bar 2e-5 / 1e-4).
constexpr int DG_T = 512;
constexpr int DG_W = DG_T / 32;
constexpr int DG_CAP = 63;
constexpr int DG_IDS = 352;

__device__ __forceinline__ float2 h0_pair(float x0, float w0a, float w0b, float w1a, float w1b) {  // embed_init_kernel's row
  const float a = fmaxf(fmaf(x0, w1a, x0 * w0a), 0.f);
  const float b = fmaxf(fmaf(x0, w1b, x0 * w0b), 0.f);
  const float den = fmaxf(sqrtf(warp_sum(a * a + b * b)), 1e-12f);
  return make_float2(a / den, b / den);
}
// columns (2 lane, 2 lane + 1) of H1(d) = normalise(relu(d a + b)); a = c @ W13, b = c @ W23 (this lane's two columns each)
__device__ __forceinline__ float2 h1_pair(float d, float2 a, float2 b) {
  const float x = d > 0.f ? fmaxf(fmaf(d, a.x, b.x), 0.f) : 0.f;  // (d = 0: isolated node, first embedding 0, no neighbours)
  const float y = d > 0.f ? fmaxf(fmaf(d, a.y, b.y), 0.f) : 0.f;
  const float den = fmaxf(sqrtf(warp_sum(x * x + y * y)), 1e-12f);
  return make_float2(x / den, y / den);
}
// c (shared, 64 floats) and this lane's columns of a = c @ W13, b = c @ W23; all threads of the CTA call it
__device__ __forceinline__ void h1_coeffs(const md_weights_t& W, float* c_sm, float2& a, float2& b) {
  const int lane = threadIdx.x & 31;
  if (threadIdx.x < 32) {
    const float2 c = h0_pair(1.f, W.w_n2l[2 * lane], W.w_n2l[2 * lane + 1], W.w_n2l[D + 2 * lane], W.w_n2l[D + 2 * lane + 1]);
    c_sm[2 * lane] = c.x, c_sm[2 * lane + 1] = c.y;
  }
  __syncthreads();
  a = make_float2(0.f, 0.f), b = make_float2(0.f, 0.f);
  for (int k = 0; k < D; ++k) {
    const float ck = c_sm[k];
    const float2 w13 = *reinterpret_cast<const float2*>(W.w13 + k * D + 2 * lane);
    const float2 w23 = *reinterpret_cast<const float2*>(W.w23 + k * D + 2 * lane);
    a.x = fmaf(ck, w13.x, a.x), a.y = fmaf(ck, w13.y, a.y);
    b.x = fmaf(ck, w23.x, b.x), b.y = fmaf(ck, w23.y, b.y);
  }
}

template <int LEVEL>
__global__ void __launch_bounds__(DG_T, 2) agg_deg_kernel(md_graphs_t G, md_state_t S, md_weights_t W, md_work_t K) {
  __shared__ float2 tab[DG_CAP + 1][32];
  __shared__ int du[DG_W][DG_IDS];
  __shared__ float c_sm[D];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int B = G.n_graphs, St = G.stride;
  const int l = blockIdx.y / B, g = blockIdx.y - l * B;
  const int n = S.n_act[g];
  if (n <= 0) return;
  const float w0a = W.w_n2l[2 * lane], w0b = W.w_n2l[2 * lane + 1];
  const float w1a = W.w_n2l[D + 2 * lane], w1b = W.w_n2l[D + 2 * lane + 1];
  const float dmax = (float)S.max_deg[l * B + g];
  float2 ca, cb;
  if constexpr (LEVEL == 1) h1_coeffs(W, c_sm, ca, cb);
  auto row_of = [&](int d) -> float2 {  // this lane's two columns of the table row of alive degree d (whole warp calls it)
    if constexpr (LEVEL == 1) return h1_pair((float)d, ca, cb);
    else return h0_pair((float)d / dmax, w0a, w0b, w1a, w1b);
  };
  for (int d = warp; d <= DG_CAP; d += DG_W) tab[d][lane] = row_of(d);
  __syncthreads();
  const int32_t* __restrict__ deg = S.deg[l];
  const int32_t* __restrict__ acol_l = S.acol[l];
  // iteration 0 writes its neighbour sums into buffer 1, iteration 1 into buffer 0 (where the transform expects them)
  float2* __restrict__ Pout = reinterpret_cast<float2*>(K.h[1 - LEVEL][l]) + lane;
  const int n_chunks = (n + 31) / 32;
  for (int c = blockIdx.x * DG_W + warp; c < n_chunks; c += gridDim.x * DG_W) {
    const int p0 = 32 * c, rows = min(32, n - p0);
    const int32_t* __restrict__ aptr = S.aptr[l] + (size_t)g * St + p0;
    const int ap = aptr[min(lane, rows)], ap_hi = aptr[min(lane + 1, rows)];
    const int v = lane < rows ? S.act_list[(size_t)g * St + p0 + lane] : -1;
    const int e_begin = __shfl_sync(0xffffffffu, ap, 0);
    const int total = __shfl_sync(0xffffffffu, ap_hi, 31) - e_begin;
    const int32_t* __restrict__ acol = acol_l + e_begin;
    bool simple = total <= DG_IDS;
    for (int k = lane; k < min(total, DG_IDS); k += 32) {
      const int d = deg[acol[k]];
      du[warp][k] = d;
      simple &= d <= DG_CAP;
    }
    simple = __all_sync(0xffffffffu, simple);
    __syncwarp();
    if (simple) {
      // A quarter-warp per destination row (8 lanes x 32 bytes), two rows per quarter at a time: per table row added the
      // warp issues 1.75 instructions instead of 8 (a whole warp per 256-byte row was issue-bound: 0.76 ms per launch).
      const int q = lane >> 3, j8 = lane & 7;
      // lane j8 owns columns [4 j8, 4 j8 + 4) and [32 + 4 j8, 36 + 4 j8): every 8-lane phase of an LDS.128 / STG.128 then
      // covers 128 contiguous bytes (no bank conflicts, full sectors)
      const float4* tab4 = reinterpret_cast<const float4*>(&tab[0][0]) + j8;  // row d: tab4[16 * d], tab4[16 * d + 8]
      float4* __restrict__ Pout4 = reinterpret_cast<float4*>(K.h[1 - LEVEL][l]) + j8;
      for (int pass = 0; pass < 4 && 8 * pass < rows; ++pass) {
        const int rA = 8 * pass + q, rB = rA + 4;
        const int e0A = __shfl_sync(0xffffffffu, ap, rA) - e_begin, e0B = __shfl_sync(0xffffffffu, ap, rB) - e_begin;
        const int e1A = __shfl_sync(0xffffffffu, ap_hi, rA) - e_begin, e1B = __shfl_sync(0xffffffffu, ap_hi, rB) - e_begin;
        const int lenA = rA < rows ? e1A - e0A : 0, lenB = rB < rows ? e1B - e0B : 0;  // (shuffles stay outside the selects)
        const int vA = __shfl_sync(0xffffffffu, v, rA), vB = __shfl_sync(0xffffffffu, v, rB);
        int maxlen = max(lenA, lenB);
        maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, 8));
        maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, 16));
        float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, b0 = a0, b1 = a0;
        for (int i = 0; i < maxlen; ++i) {
          if (i < lenA) {
            const float4* t = tab4 + 16 * du[warp][e0A + i];
            const float4 x0 = t[0], x1 = t[8];
            a0.x += x0.x, a0.y += x0.y, a0.z += x0.z, a0.w += x0.w;
            a1.x += x1.x, a1.y += x1.y, a1.z += x1.z, a1.w += x1.w;
          }
          if (i < lenB) {
            const float4* t = tab4 + 16 * du[warp][e0B + i];
            const float4 x0 = t[0], x1 = t[8];
            b0.x += x0.x, b0.y += x0.y, b0.z += x0.z, b0.w += x0.w;
            b1.x += x1.x, b1.y += x1.y, b1.z += x1.z, b1.w += x1.w;
          }
        }
        if (rA < rows) Pout4[(size_t)vA * (D / 4)] = a0, Pout4[(size_t)vA * (D / 4) + 8] = a1;
        if (rB < rows) Pout4[(size_t)vB * (D / 4)] = b0, Pout4[(size_t)vB * (D / 4) + 8] = b1;
      }
    } else {  // hub rows beyond the staging area or degrees beyond the table
      for (int r = 0; r < rows; ++r) {
        const int e0 = __shfl_sync(0xffffffffu, ap, r) - e_begin, e1 = __shfl_sync(0xffffffffu, ap_hi, r) - e_begin;
        const int vv = __shfl_sync(0xffffffffu, v, r);
        float2 acc = make_float2(0.f, 0.f);
        for (int e = e0; e < e1; ++e) {
          const int d = e < DG_IDS ? du[warp][e] : deg[acol[e]];  // (warp-uniform)
          const float2 x = d <= DG_CAP ? tab[d][lane] : row_of(d);
          acc.x += x.x;
          acc.y += x.y;
        }
        Pout[(size_t)vv * (D / 2)] = acc;
      }
    }
    __syncwarp();
  }
}

}  // namespace md

extern int g_md_debug_flags;

// enqueue P = A_adj . H for every layer (internal: called by md_embed_iter when the two-kernel path applies)
int md_agg_rows_launch(const md_graphs_t* g, const md_state_t* s, const md_work_t* k, int iter, cudaStream_t st) {
  static md::PerDevice<int> nsm_pd;
  int& nsm = nsm_pd.get();
  if (nsm <= 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    if (nsm <= 0) nsm = 148;
  }
  const long long n_chunks = (long long)g->n_graphs * (g->stride / 32) * g->n_layers;
  if ((g_md_debug_flags & 512) || g->n_total >= (1 << md::RG_IDBITS)) {  // (the ring packs the row into the id's top bits)
    const long long want = (n_chunks + md::AG_W - 1) / md::AG_W;
    const int grid = (int)(want < nsm ? want : nsm);
    md::agg_rows_kernel<<<grid, md::AG_T, 0, st>>>(*g, *s, *k, iter);
  } else {
    static md::PerDevice<bool> configured_pd;
  bool& configured = configured_pd.get();
    const int smem = (int)sizeof(md::RingSmem);
    if (!configured) {
      cudaError_t e = cudaFuncSetAttribute(md::agg_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      if (e != cudaSuccess) return (int)e;
      configured = true;
    }
    const long long want = (n_chunks + md::RG_W - 1) / md::RG_W;
    const int grid = (int)(want < nsm ? want : nsm);
    md::agg_ring_kernel<<<grid, md::RG_T, smem, st>>>(*g, *s, *k, iter);
  }
  MD_CHECK_LAUNCH();
  return 0;
}

// iteration 0 (level 0) / iteration 1 (level 1) of the unit-cost variant on the two-kernel path (see agg_deg_kernel)
int md_agg_deg_launch(const md_graphs_t* g, const md_state_t* s, const md_weights_t* w, const md_work_t* k, int level, cudaStream_t st) {
  static md::PerDevice<int> nsm_pd;
  int& nsm = nsm_pd.get();
  if (nsm <= 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    if (nsm <= 0) nsm = 148;
  }
  const int units = g->n_graphs * g->n_layers;
  int per = 2 * nsm / units;  // CTAs per (layer, graph): one wave of two CTAs per SM
  const int max_per = (g->stride / 32 + md::DG_W - 1) / md::DG_W;
  if (per > max_per) per = max_per;
  if (per < 1) per = 1;
  if (level == 1) md::agg_deg_kernel<1><<<dim3(per, units), md::DG_T, 0, st>>>(*g, *s, *w, *k);
  else md::agg_deg_kernel<0><<<dim3(per, units), md::DG_T, 0, st>>>(*g, *s, *w, *k);
  MD_CHECK_LAUNCH();
  return 0;
}
