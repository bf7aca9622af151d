// Dense primitives of the DQN training step (forward AND backward), FP32, hand-written: every arithmetic op of
// train_forward / calc_loss / backward runs in one of these kernels (or in md_spmm_csr_f32); torch autograd only chains them.
//
// Replaces (reference, /root/reference/code/MultiDismantler_unit_cost):
//   MultiDismantler_net_graphsage.py:92-241   train_forward: relu(x @ w_n2l), F.normalize, the three dense transforms
//                                             (p_node_conv, p_node_conv2, p_node_conv3), the Q head (cross_product, h1_weight,
//                                             h2_weight / last_w), the layer gate (w_layer1, w_layer2) and its softmax mix
//   MRGNN/mutil_layer_weight.py:266-313       BitwiseMultipyLogis: tanh(E @ trans + bias), sigmoid-of-dot gates, softmax, mix
//   MultiDismantler_torch.py:410-431          calc_loss: MSE + Laplacian reconstruction term (with md_spmm_csr_f32)
// and the backward of each of them (the reference gets it from torch autograd).
//
// Sizes are small (n <= ~13 k rows, widths <= 128, 31 k parameters): plain FP32 SIMT tiles, deterministic reductions
// (fixed order, no floating-point atomics), no tensor cores -- the training step is host / launch bound, not FLOP bound.
#include "md_common.cuh"

namespace md {

// ---- C = [C +] op(A) op(B) ---------------------------------------------------------------------------------------
// A is [M, K] row-major (transA: stored [K, M]); B is [K, N] row-major (transB: stored [N, K]); relu_a clamps A at 0 on load.
constexpr int GM_T = 256, GM_BM = 64, GM_BN = 64, GM_BK = 16;

// Split K (gridDim.z > 1): slice z covers k in [z * kchunk, min(K, (z + 1) * kchunk)) and writes its partial tile to
// C + z * M * N (the caller points C at a workspace; splitk_reduce_kernel folds the slices in a fixed order).
__global__ void __launch_bounds__(GM_T) gemm_f32_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C,
                                                       int M, int N, int K, int transA, int transB, int accumulate, int relu_a, int kchunk) {
  __shared__ float As[GM_BK][GM_BM + 1], Bs[GM_BK][GM_BN + 1];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * GM_BM, n0 = blockIdx.x * GM_BN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int kbeg = blockIdx.z * kchunk, kend = min(K, kbeg + kchunk);
  C += (size_t)blockIdx.z * M * N;
  for (int k0 = kbeg; k0 < kend; k0 += GM_BK) {
    for (int i = tid; i < GM_BK * GM_BM; i += GM_T) {
      int kk, mm;
      if (transA) {  // stored [K, M]: consecutive threads walk m
        kk = i / GM_BM;
        mm = i - kk * GM_BM;
      } else {  // stored [M, K]: consecutive threads walk k
        mm = i / GM_BK;
        kk = i - mm * GM_BK;
      }
      const int m = m0 + mm, k = k0 + kk;
      float v = 0.f;
      if (m < M && k < kend) v = transA ? A[(size_t)k * M + m] : A[(size_t)m * K + k];
      if (relu_a) v = fmaxf(v, 0.f);
      As[kk][mm] = v;
    }
    for (int i = tid; i < GM_BK * GM_BN; i += GM_T) {
      int kk, nn;
      if (transB) {  // stored [N, K]
        nn = i / GM_BK;
        kk = i - nn * GM_BK;
      } else {  // stored [K, N]
        kk = i / GM_BN;
        nn = i - kk * GM_BN;
      }
      const int n = n0 + nn, k = k0 + kk;
      float v = 0.f;
      if (n < N && k < kend) v = transB ? B[(size_t)n * K + k] : B[(size_t)k * N + n];
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GM_BK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty + 16 * i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx + 16 * j;
      if (n < N) {
        float* c = C + (size_t)m * N + n;
        *c = accumulate ? *c + acc[i][j] : acc[i][j];
      }
    }
  }
}

// C[i] = [C[i] +] sum_z P[z][i]   (z ascending: deterministic)
__global__ void splitk_reduce_kernel(const float* __restrict__ P, float* __restrict__ C, int S, int MN, int accumulate) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < MN; i += gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int z = 0; z < S; ++z) s += P[(size_t)z * MN + i];
    C[i] = accumulate ? C[i] + s : s;
  }
}

// ---- row-wise L2 normalise (optionally of relu(Z)) : H = x / max(|x|, 1e-12), x = relu?(Z) ------------------------------
// one warp per row, D <= 128 (4 values per lane)
__global__ void __launch_bounds__(256) norm_fwd_kernel(const float* __restrict__ Z, float* __restrict__ H, int n, int D, int relu) {
  const int lane = threadIdx.x & 31;
  for (int r = blockIdx.x * 8 + (threadIdx.x >> 5); r < n; r += gridDim.x * 8) {
    float x[4], ss = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = lane + 32 * j;
      x[j] = c < D ? Z[(size_t)r * D + c] : 0.f;
      if (relu) x[j] = fmaxf(x[j], 0.f);
      ss = fmaf(x[j], x[j], ss);
    }
    ss = warp_sum(ss);
    const float den = fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = lane + 32 * j;
      if (c < D) H[(size_t)r * D + c] = x[j] / den;
    }
  }
}
// dZ = relu'(Z) * (dH - H (H . dH)) / max(|x|, eps)     (|x| recomputed from Z; for |x| <= eps: dZ = relu'(Z) dH / eps)
__global__ void __launch_bounds__(256) norm_bwd_kernel(const float* __restrict__ Z, const float* __restrict__ dH, float* __restrict__ dZ,
                                                      int n, int D, int relu) {
  const int lane = threadIdx.x & 31;
  for (int r = blockIdx.x * 8 + (threadIdx.x >> 5); r < n; r += gridDim.x * 8) {
    float x[4], g[4], ss = 0.f, dot = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = lane + 32 * j;
      x[j] = c < D ? Z[(size_t)r * D + c] : 0.f;
      g[j] = c < D ? dH[(size_t)r * D + c] : 0.f;
      if (relu) x[j] = fmaxf(x[j], 0.f);
      ss = fmaf(x[j], x[j], ss);
      dot = fmaf(x[j], g[j], dot);
    }
    ss = warp_sum(ss);
    dot = warp_sum(dot);
    const float nrm = sqrtf(ss);
    const bool clamped = nrm <= 1e-12f;
    const float den = fmaxf(nrm, 1e-12f);
    const float k = clamped ? 0.f : dot / (den * den * den);  // d/dx (x / |x|) = I / |x| - x x^T / |x|^3
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = lane + 32 * j;
      if (c < D) {
        float d = g[j] / den - x[j] * k;
        if (relu && !(Z[(size_t)r * D + c] > 0.f)) d = 0.f;
        dZ[(size_t)r * D + c] = d;
      }
    }
  }
}

// ---- T = tanh(Z + b) ; dZ = dT (1 - T^2) -----------------------------------------------------------------------------
__global__ void tanh_bias_fwd_kernel(const float* __restrict__ Z, const float* __restrict__ b, float* __restrict__ T, long long total, int D) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    T[i] = tanhf(Z[i] + b[i % D]);
}
__global__ void tanh_bwd_kernel(const float* __restrict__ T, const float* __restrict__ dT, float* __restrict__ dZ, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const float t = T[i];
    dZ[i] = dT[i] * (1.f - t * t);
  }
}

// ---- column sums: out[c] = sum_r X[r, c]   (fixed order: 8 row groups per block column, then a tree over the groups) ------
// gridDim.y > 1: block y sums the rows [y * rchunk, min(n, (y + 1) * rchunk)) into out + y * D (folded by splitk_reduce_kernel)
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ X, float* __restrict__ out, int n, int D, int rchunk) {
  __shared__ float part[8][32];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), q = threadIdx.x >> 5;
  const int rbeg = blockIdx.y * rchunk, rend = min(n, rbeg + rchunk);
  out += (size_t)blockIdx.y * D;
  float s = 0.f;
  if (c < D)
    for (int r = rbeg + q; r < rend; r += 8) s += X[(size_t)r * D + c];
  part[q][threadIdx.x & 31] = s;
  __syncthreads();
  if (q == 0 && c < D) {
    const int l = threadIdx.x & 31;
    out[c] = ((part[0][l] + part[1][l]) + (part[2][l] + part[3][l])) + ((part[4][l] + part[5][l]) + (part[6][l] + part[7][l]));
  }
}

// ---- inter-layer fusion of two layers (BitwiseMultipyLogis, mutil_layer_weight.py:266-301) ----------------------------------
//   a[p][k] = sigmoid(sum_c T_k T_p lw + lb);  alpha[p] = softmax_k(a[p][.]);  X_p = T_p + alpha[p][1-p] T_{1-p}
// one warp per row, 2 columns per lane.  Saves alpha-related scalars for the backward: A[r] = {a00, a01, a11, w0, w1} with
// w_p = alpha[p][1-p].
__global__ void __launch_bounds__(256) fuse2_fwd_kernel(const float* __restrict__ T0, const float* __restrict__ T1, const float* __restrict__ lw,
                                                       const float* __restrict__ lb, float* __restrict__ X0, float* __restrict__ X1,
                                                       float* __restrict__ Asv, int n) {
  const int lane = threadIdx.x & 31;
  const float2 w = *reinterpret_cast<const float2*>(lw + 2 * lane);
  const float b = lb[0];
  for (int r = blockIdx.x * 8 + (threadIdx.x >> 5); r < n; r += gridDim.x * 8) {
    const float2 t0 = *reinterpret_cast<const float2*>(T0 + (size_t)r * D + 2 * lane);
    const float2 t1 = *reinterpret_cast<const float2*>(T1 + (size_t)r * D + 2 * lane);
    const float d00 = warp_sum(fmaf(t0.y * t0.y, w.y, t0.x * t0.x * w.x));
    const float d01 = warp_sum(fmaf(t0.y * t1.y, w.y, t0.x * t1.x * w.x));
    const float d11 = warp_sum(fmaf(t1.y * t1.y, w.y, t1.x * t1.x * w.x));
    const float a00 = 1.f / (1.f + expf(-(d00 + b))), a01 = 1.f / (1.f + expf(-(d01 + b))), a11 = 1.f / (1.f + expf(-(d11 + b)));
    // p = 0: softmax over (a00, a01), weight of layer 1;  p = 1: softmax over (a01, a11), weight of layer 0
    const float m0 = fmaxf(a00, a01), e00 = expf(a00 - m0), e01 = expf(a01 - m0), w0 = e01 / (e00 + e01);
    const float m1 = fmaxf(a01, a11), f01 = expf(a01 - m1), f11 = expf(a11 - m1), w1 = f01 / (f01 + f11);
    *reinterpret_cast<float2*>(X0 + (size_t)r * D + 2 * lane) = make_float2(fmaf(w0, t1.x, t0.x), fmaf(w0, t1.y, t0.y));
    *reinterpret_cast<float2*>(X1 + (size_t)r * D + 2 * lane) = make_float2(fmaf(w1, t0.x, t1.x), fmaf(w1, t0.y, t1.y));
    if (lane == 0) {
      float* s = Asv + (size_t)r * 8;
      s[0] = a00, s[1] = a01, s[2] = a11, s[3] = w0, s[4] = w1;
    }
  }
}
// backward: dT0, dT1 [n, 64]; per-row contributions Glw [n, 64] (d lw) and Glb [n] (d lb), column-summed by the caller
__global__ void __launch_bounds__(256) fuse2_bwd_kernel(const float* __restrict__ T0, const float* __restrict__ T1, const float* __restrict__ lw,
                                                       const float* __restrict__ Asv, const float* __restrict__ dX0,
                                                       const float* __restrict__ dX1, float* __restrict__ dT0, float* __restrict__ dT1,
                                                       float* __restrict__ Glw, float* __restrict__ Glb, int n) {
  const int lane = threadIdx.x & 31;
  const float2 w = *reinterpret_cast<const float2*>(lw + 2 * lane);
  for (int r = blockIdx.x * 8 + (threadIdx.x >> 5); r < n; r += gridDim.x * 8) {
    const float2 t0 = *reinterpret_cast<const float2*>(T0 + (size_t)r * D + 2 * lane);
    const float2 t1 = *reinterpret_cast<const float2*>(T1 + (size_t)r * D + 2 * lane);
    const float2 g0 = *reinterpret_cast<const float2*>(dX0 + (size_t)r * D + 2 * lane);
    const float2 g1 = *reinterpret_cast<const float2*>(dX1 + (size_t)r * D + 2 * lane);
    const float* s = Asv + (size_t)r * 8;
    const float a00 = s[0], a01 = s[1], a11 = s[2], w0 = s[3], w1 = s[4];
    // X0 = T0 + w0 T1, X1 = T1 + w1 T0
    const float dw0 = warp_sum(fmaf(g0.y, t1.y, g0.x * t1.x));  // dL/dw0
    const float dw1 = warp_sum(fmaf(g1.y, t0.y, g1.x * t0.x));
    // w0 = softmax(a00, a01)[1]: dw0/da01 = w0 (1 - w0), dw0/da00 = -w0 (1 - w0);  w1 = softmax(a01, a11)[0]: dw1/da01 = w1 (1 - w1), dw1/da11 = -w1 (1 - w1)
    const float da00 = -dw0 * w0 * (1.f - w0);
    const float da01 = dw0 * w0 * (1.f - w0) + dw1 * w1 * (1.f - w1);
    const float da11 = -dw1 * w1 * (1.f - w1);
    // a = sigmoid(d + lb): dd = da * a (1 - a)
    const float dd00 = da00 * a00 * (1.f - a00), dd01 = da01 * a01 * (1.f - a01), dd11 = da11 * a11 * (1.f - a11);
    // d00 = sum T0 T0 lw, d01 = sum T0 T1 lw, d11 = sum T1 T1 lw
    float2 o0, o1, gl;
    o0.x = g0.x + w1 * g1.x + (2.f * dd00 * t0.x + dd01 * t1.x) * w.x;
    o0.y = g0.y + w1 * g1.y + (2.f * dd00 * t0.y + dd01 * t1.y) * w.y;
    o1.x = g1.x + w0 * g0.x + (2.f * dd11 * t1.x + dd01 * t0.x) * w.x;
    o1.y = g1.y + w0 * g0.y + (2.f * dd11 * t1.y + dd01 * t0.y) * w.y;
    gl.x = dd00 * t0.x * t0.x + dd01 * t0.x * t1.x + dd11 * t1.x * t1.x;
    gl.y = dd00 * t0.y * t0.y + dd01 * t0.y * t1.y + dd11 * t1.y * t1.y;
    *reinterpret_cast<float2*>(dT0 + (size_t)r * D + 2 * lane) = o0;
    *reinterpret_cast<float2*>(dT1 + (size_t)r * D + 2 * lane) = o1;
    *reinterpret_cast<float2*>(Glw + (size_t)r * D + 2 * lane) = gl;
    if (lane == 0) Glb[r] = dd00 + dd01 + dd11;
  }
}

// ---- rows scaled by a per-graph scalar: E[r, :] = H[r, :] * s[gid[r]] ; dH = dE * s[gid] ; rowdot[r] = dE[r] . H[r] -----
__global__ void rowscale_kernel(const float* __restrict__ H, const float* __restrict__ s, const int32_t* __restrict__ gid, float* __restrict__ E,
                                long long total, int D) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    E[i] = H[i] * s[gid[i / D]];
}
__global__ void __launch_bounds__(256) rowdot_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ out, int n, int D) {
  const int lane = threadIdx.x & 31;
  for (int r = blockIdx.x * 8 + (threadIdx.x >> 5); r < n; r += gridDim.x * 8) {
    float s = 0.f;
    for (int c = lane; c < D; c += 32) s = fmaf(A[(size_t)r * D + c], B[(size_t)r * D + c], s);
    s = warp_sum(s);
    if (lane == 0) out[r] = s;
  }
}
// out[g] = sum_{r in [off[g], off[g+1])} x[r] * (y ? y[r] : 1): one warp per graph, fixed order
__global__ void __launch_bounds__(256) seg_reduce_kernel(const float* __restrict__ x, const float* __restrict__ y, const int64_t* __restrict__ off,
                                                        float* __restrict__ out, int B) {
  const int lane = threadIdx.x & 31;
  for (int g = blockIdx.x * 8 + (threadIdx.x >> 5); g < B; g += gridDim.x * 8) {
    float s = 0.f;
    for (long long r = off[g] + lane; r < off[g + 1]; r += 32) s = fmaf(x[r], y ? y[r] : 1.f, s);
    s = warp_sum(s);
    if (lane == 0) out[g] = s;
  }
}

// ---- layer-gate mix of the two per-layer Q values ------------------------------------------------------------------------
//   sm = softmax(gate0[g], gate1[g]);  q[r] = sm0 (q0[r] + a0[g]) + sm1 (q1[r] + a1[g])       g = gid[r]
// q_l[r] = relu(hidden_l[r]) . h2[:32] is per node, a_l[g] = aux_l[g] . h2[32:] per graph (net_graphsage.py:214-241: the
// reference concatenates [hidden, aux] and multiplies by last_w once).
__device__ __forceinline__ float gate_sm0(float a, float b) {
  const float m = fmaxf(a, b), ea = expf(a - m), eb = expf(b - m);
  return ea / (ea + eb);
}
__global__ void head_mix_fwd_kernel(const float* __restrict__ g0, const float* __restrict__ g1, const float* __restrict__ a0,
                                    const float* __restrict__ a1, const float* __restrict__ q0, const float* __restrict__ q1,
                                    const int32_t* __restrict__ gid, float* __restrict__ q, int n) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
    const int g = gid[r];
    const float s0 = gate_sm0(g0[g], g1[g]);
    q[r] = s0 * (q0[r] + a0[g]) + (1.f - s0) * (q1[r] + a1[g]);
  }
}
// rows: dq0 = sm0 dq, dq1 = sm1 dq
__global__ void head_mix_bwd_rows_kernel(const float* __restrict__ g0, const float* __restrict__ g1, const float* __restrict__ dq,
                                         const int32_t* __restrict__ gid, float* __restrict__ dq0, float* __restrict__ dq1, int n) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
    const int g = gid[r];
    const float s0 = gate_sm0(g0[g], g1[g]);
    dq0[r] = s0 * dq[r];
    dq1[r] = (1.f - s0) * dq[r];
  }
}
// graphs: with dsum_q[g] = sum_r dq (q0 - q1) and dqs[g] = sum_r dq over the rows of g:
//   dgate0 = sm0 sm1 (dsum_q + dqs (a0 - a1)), dgate1 = -dgate0, da0 = sm0 dqs, da1 = sm1 dqs
__global__ void head_mix_bwd_graph_kernel(const float* __restrict__ g0, const float* __restrict__ g1, const float* __restrict__ a0,
                                          const float* __restrict__ a1, const float* __restrict__ dsum_q, const float* __restrict__ dqs,
                                          float* __restrict__ dg0, float* __restrict__ dg1, float* __restrict__ da0,
                                          float* __restrict__ da1, int B) {
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < B; g += gridDim.x * blockDim.x) {
    const float s0 = gate_sm0(g0[g], g1[g]);
    const float d = s0 * (1.f - s0) * (dsum_q[g] + dqs[g] * (a0[g] - a1[g]));
    dg0[g] = d;
    dg1[g] = -d;
    da0[g] = s0 * dqs[g];
    da1[g] = (1.f - s0) * dqs[g];
  }
}

// ---- small elementwise helpers ---------------------------------------------------------------------------------------------
// out = alpha * a + beta * b   (b may be NULL)
__global__ void axpby_kernel(const float* __restrict__ a, const float* __restrict__ b, float alpha, float beta, float* __restrict__ out, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = alpha * a[i] + (b ? beta * b[i] : 0.f);
}
// dX *= (X > 0)
__global__ void relu_mask_kernel(const float* __restrict__ X, float* __restrict__ dX, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    if (!(X[i] > 0.f)) dX[i] = 0.f;
}
// out[0] = sum a[i] b[i]  : one block, fixed order
__global__ void __launch_bounds__(1024) dot_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, long long n) {
  __shared__ float part[32];
  float s = 0.f;
  for (long long i = threadIdx.x; i < n; i += 1024) s = fmaf(a[i], b[i], s);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = part[threadIdx.x];
    t = warp_sum(t);
    if (threadIdx.x == 0) out[0] = t;
  }
}

// ---- Adam (torch.optim.Adam defaults as the reference uses them, MultiDismantler_torch.py:136: no weight decay, no amsgrad) --
// bc1 = 1 - beta1^t, bc2 = 1 - beta2^t for the step being taken
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n,
                            float lr, float b1, float b2, float eps, float bc1, float bc2) {
  const float step = lr / bc1, rs2 = 1.f / sqrtf(bc2);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i];
    const float mi = m[i] + (gi - m[i]) * (1.f - b1);
    const float vi = v[i] * b2 + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= step * (mi / (sqrtf(vi) * rs2 + eps));
  }
}

static inline int blocks_for(long long n, int per) {
  long long b = (n + per - 1) / per;
  return (int)(b < 1 ? 1 : (b > 4096 ? 4096 : b));
}

}  // namespace md

using namespace md;

extern "C" int md_gemm_f32(const float* A, const float* B, float* C, int32_t M, int32_t N, int32_t K, int32_t transA, int32_t transB,
                           int32_t accumulate, int32_t relu_a, float* ws, int64_t ws_floats, md_stream_t stream) {
  if (!A || !B || !C || M < 0 || N < 0 || K < 0) return MD_EINVAL;
  if (M == 0 || N == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((N + GM_BN - 1) / GM_BN, (M + GM_BM - 1) / GM_BM);
  // long reductions into a small output (the weight gradients: K = rows of the batch): split K over up to 64 slices
  int S = 1;
  if (ws && K >= 1024 && (long long)grid.x * grid.y < 64) {
    S = (K + 255) / 256;
    if (S > 64) S = 64;
    while (S > 1 && (long long)S * M * N > ws_floats) --S;
  }
  if (S > 1) {
    int kchunk = (K + S - 1) / S;
    kchunk = (kchunk + GM_BK - 1) / GM_BK * GM_BK;
    S = (K + kchunk - 1) / kchunk;
    grid.z = S;
    gemm_f32_kernel<<<grid, GM_T, 0, st>>>(A, B, ws, M, N, K, transA, transB, 0, relu_a, kchunk);
    splitk_reduce_kernel<<<blocks_for((long long)M * N, 256), 256, 0, st>>>(ws, C, S, M * N, accumulate);
  } else {
    gemm_f32_kernel<<<grid, GM_T, 0, st>>>(A, B, C, M, N, K, transA, transB, accumulate, relu_a, K > 0 ? K : 1);
  }
  MD_CHECK_LAUNCH();
  return 0;
}
extern "C" int md_norm_fwd_f32(const float* Z, float* H, int32_t n, int32_t D, int32_t relu, md_stream_t stream) {
  if (!Z || !H || D < 1 || D > 128) return MD_EINVAL;
  if (n == 0) return 0;
  norm_fwd_kernel<<<blocks_for(n, 8), 256, 0, (cudaStream_t)stream>>>(Z, H, n, D, relu);
  MD_CHECK_LAUNCH();
  return 0;
}
extern "C" int md_norm_bwd_f32(const float* Z, const float* dH, float* dZ, int32_t n, int32_t D, int32_t relu, md_stream_t stream) {
  if (!Z || !dH || !dZ || D < 1 || D > 128) return MD_EINVAL;
  if (n == 0) return 0;
  norm_bwd_kernel<<<blocks_for(n, 8), 256, 0, (cudaStream_t)stream>>>(Z, dH, dZ, n, D, relu);
  MD_CHECK_LAUNCH();
  return 0;
}
extern "C" int md_tanh_bias_fwd_f32(const float* Z, const float* b, float* T, int32_t n, int32_t D, md_stream_t stream) {
  if (!Z || !b || !T || D < 1) return MD_EINVAL;
  if (n == 0) return 0;
  tanh_bias_fwd_kernel<<<blocks_for((long long)n * D, 256), 256, 0, (cudaStream_t)stream>>>(Z, b, T, (long long)n * D, D);
  MD_CHECK_LAUNCH();
  return 0;
}
extern "C" int md_tanh_bwd_f32(const float* T, const float* dT, float* dZ, int32_t n, int32_t D, md_stream_t stream) {
  if (!T || !dT || !dZ) return MD_EINVAL;
  if (n == 0) return 0;
  tanh_bwd_kernel<<<blocks_for((long long)n * D, 256), 256, 0, (cudaStream_t)stream>>>(T, dT, dZ, (long long)n * D);
  MD_CHECK_LAUNCH();
  return 0;
}
extern "C" int md_colsum_f32(const float* X, float* out, int32_t n, int32_t D, float* ws, int64_t ws_floats, md_stream_t stream) {
  if (!X || !out || D < 1) return MD_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  int S = 1;
  if (ws && n >= 1024) {  // many rows: row chunks in parallel, folded in chunk order
    S = (n + 255) / 256;
    if (S > 128) S = 128;
    while (S > 1 && (long long)S * D > ws_floats) --S;
  }
  if (S > 1) {
    const int rchunk = (n + S - 1) / S;
    S = (n + rchunk - 1) / rchunk;
    colsum_kernel<<<dim3((D + 31) / 32, S), 256, 0, st>>>(X, ws, n, D, rchunk);
    splitk_reduce_kernel<<<blocks_for(D, 256), 256, 0, st>>>(ws, out, S, D, 0);
  } else {
    colsum_kernel<<<(D + 31) / 32, 256, 0, st>>>(X, out, n, D, n > 0 ? n : 1);
  }
  MD_CHECK_LAUNCH();
  return 0;
}
extern "C" int md_fuse2_fwd_f32(const float* T0, const float* T1, const float* lw, const float* lb, float* X0, float* X1, float* saved,
                                int32_t n, md_stream_t stream) {
  if (!T0 || !T1 || !lw || !lb || !X0 || !X1 || !saved) return MD_EINVAL;
  if (n == 0) return 0;
  fuse2_fwd_kernel<<<blocks_for(n, 8), 256, 0, (cudaStream_t)stream>>>(T0, T1, lw, lb, X0, X1, saved, n);
  MD_CHECK_LAUNCH();
  return 0;
}
extern "C" int md_fuse2_bwd_f32(const float* T0, const float* T1, const float* lw, const float* saved, const float* dX0, const float* dX1,
                                float* dT0, float* dT1, float* Glw, float* Glb, int32_t n, md_stream_t stream) {
  if (!T0 || !T1 || !lw || !saved || !dX0 || !dX1 || !dT0 || !dT1 || !Glw || !Glb) return MD_EINVAL;
  if (n == 0) return 0;
  fuse2_bwd_kernel<<<blocks_for(n, 8), 256, 0, (cudaStream_t)stream>>>(T0, T1, lw, saved, dX0, dX1, dT0, dT1, Glw, Glb, n);
  MD_CHECK_LAUNCH();
  return 0;
}
extern "C" int md_rowscale_f32(const float* H, const float* s, const int32_t* gid, float* E, int32_t n, int32_t D, md_stream_t stream) {
  if (!H || !s || !gid || !E) return MD_EINVAL;
  if (n == 0) return 0;
  rowscale_kernel<<<blocks_for((long long)n * D, 256), 256, 0, (cudaStream_t)stream>>>(H, s, gid, E, (long long)n * D, D);
  MD_CHECK_LAUNCH();
  return 0;
}
extern "C" int md_rowdot_f32(const float* A, const float* B, float* out, int32_t n, int32_t D, md_stream_t stream) {
  if (!A || !B || !out) return MD_EINVAL;
  if (n == 0) return 0;
  rowdot_kernel<<<blocks_for(n, 8), 256, 0, (cudaStream_t)stream>>>(A, B, out, n, D);
  MD_CHECK_LAUNCH();
  return 0;
}
extern "C" int md_seg_reduce_f32(const float* x, const float* y, const int64_t* off, float* out, int32_t B, md_stream_t stream) {
  if (!x || !off || !out) return MD_EINVAL;
  if (B == 0) return 0;
  seg_reduce_kernel<<<blocks_for(B, 8), 256, 0, (cudaStream_t)stream>>>(x, y, off, out, B);
  MD_CHECK_LAUNCH();
  return 0;
}
extern "C" int md_head_mix_fwd_f32(const float* g0, const float* g1, const float* a0, const float* a1, const float* q0, const float* q1,
                                   const int32_t* gid, float* q, int32_t n, md_stream_t stream) {
  if (!g0 || !g1 || !a0 || !a1 || !q0 || !q1 || !gid || !q) return MD_EINVAL;
  if (n == 0) return 0;
  head_mix_fwd_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(g0, g1, a0, a1, q0, q1, gid, q, n);
  MD_CHECK_LAUNCH();
  return 0;
}
extern "C" int md_head_mix_bwd_f32(const float* g0, const float* g1, const float* a0, const float* a1, const float* dq, const float* dsum_q,
                                   const float* dqs, const int32_t* gid, float* dq0, float* dq1, float* dg0, float* dg1, float* da0,
                                   float* da1, int32_t n, int32_t B, md_stream_t stream) {
  if (!g0 || !g1 || !a0 || !a1 || !dq || !dsum_q || !dqs || !gid || !dq0 || !dq1 || !dg0 || !dg1 || !da0 || !da1) return MD_EINVAL;
  if (n > 0) head_mix_bwd_rows_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(g0, g1, dq, gid, dq0, dq1, n);
  if (B > 0) head_mix_bwd_graph_kernel<<<blocks_for(B, 256), 256, 0, (cudaStream_t)stream>>>(g0, g1, a0, a1, dsum_q, dqs, dg0, dg1, da0, da1, B);
  MD_CHECK_LAUNCH();
  return 0;
}
extern "C" int md_axpby_f32(const float* a, const float* b, float alpha, float beta, float* out, int64_t n, md_stream_t stream) {
  if (!a || !out) return MD_EINVAL;
  if (n == 0) return 0;
  axpby_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(a, b, alpha, beta, out, n);
  MD_CHECK_LAUNCH();
  return 0;
}
extern "C" int md_relu_mask_f32(const float* X, float* dX, int64_t n, md_stream_t stream) {
  if (!X || !dX) return MD_EINVAL;
  if (n == 0) return 0;
  relu_mask_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(X, dX, n);
  MD_CHECK_LAUNCH();
  return 0;
}
extern "C" int md_dot_f32(const float* a, const float* b, float* out, int64_t n, md_stream_t stream) {
  if (!a || !b || !out) return MD_EINVAL;
  dot_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(a, b, out, n);
  MD_CHECK_LAUNCH();
  return 0;
}
extern "C" int md_adam_step_f32(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                                int32_t t, md_stream_t stream) {
  if (!p || !g || !m || !v || t < 1) return MD_EINVAL;
  if (n == 0) return 0;
  const float bc1 = (float)(1.0 - pow((double)beta1, (double)t)), bc2 = (float)(1.0 - pow((double)beta2, (double)t));
  adam_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, bc1, bc2);
  MD_CHECK_LAUNCH();
  return 0;
}
