// CSR SpMM, the operator-level boundary: drop-in for torch_sparse.spmm (torch-sparse 0.6.18) as the reference
// calls it at MultiDismantler_net_graphsage.py:139,146,199,290,297,350,377 and MultiDismantler_torch.py:420.
//   Y[r, :] = sum_{e in row r} val[e] * X[col[e], :]        (val == NULL: all ones, aggregatorID = 0)
// One warp per row (a half-warp per row for D <= 64, so that no lane idles at the embedding width); each lane owns float4
// column chunks; neighbours are taken in CSR order so the f32 sum is deterministic and matches a sequential scatter-add in
// COO order.
#include "md_common.cuh"

namespace md {

template <int VEC, int LPR>  // VEC float4 chunks per lane; LPR lanes per row (16: two rows per warp when D <= 64, else 32)
__global__ void __launch_bounds__(256) spmm_csr_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                      const float* __restrict__ val, const float* __restrict__ X,
                                                      float* __restrict__ Y, int n_rows, int D4) {
  constexpr int RPW = 32 / LPR;  // rows per warp
  const int lane = threadIdx.x & 31, sub = lane / LPR, sl = lane % LPR;
  const int wpb = blockDim.x >> 5;
  for (int row0 = (blockIdx.x * wpb + (threadIdx.x >> 5)) * RPW; row0 < n_rows; row0 += gridDim.x * wpb * RPW) {
    const int row = row0 + sub;
    const bool live = row < n_rows;
    const int b = live ? rowptr[row] : 0, e = live ? rowptr[row + 1] : 0;
    float4 acc[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int c0 = b; __any_sync(0xffffffffu, c0 < e); c0 += LPR) {  // (the rows of a warp may differ in length)
      int u = 0;
      float w = 1.f;
      if (c0 + sl < e) {
        u = col[c0 + sl];
        if (val) w = val[c0 + sl];
      }
      const int cnt = max(0, min(LPR, e - c0));
      for (int j0 = 0; __any_sync(0xffffffffu, j0 < cnt); j0 += 4) {
        float4 x[4][VEC];
        float ws[4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int srcl = sub * LPR + ((j0 + jj) % LPR);
          const int uu = __shfl_sync(0xffffffffu, u, srcl);
          ws[jj] = __shfl_sync(0xffffffffu, w, srcl);
          const bool ok = j0 + jj < cnt;
#pragma unroll
          for (int j = 0; j < VEC; ++j) {
            x[jj][j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ok && sl + LPR * j < D4) x[jj][j] = __ldg(reinterpret_cast<const float4*>(X) + (size_t)uu * D4 + sl + LPR * j);
          }
        }
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
          if (j0 + jj < cnt) {  // neighbours in CSR order: the f32 sum is deterministic
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
              acc[j].x = fmaf(ws[jj], x[jj][j].x, acc[j].x);
              acc[j].y = fmaf(ws[jj], x[jj][j].y, acc[j].y);
              acc[j].z = fmaf(ws[jj], x[jj][j].z, acc[j].z);
              acc[j].w = fmaf(ws[jj], x[jj][j].w, acc[j].w);
            }
          }
      }
    }
    if (live) {
#pragma unroll
      for (int j = 0; j < VEC; ++j)
        if (sl + LPR * j < D4) reinterpret_cast<float4*>(Y)[(size_t)row * D4 + sl + LPR * j] = acc[j];
    }
  }
}

}  // namespace md

extern "C" int md_spmm_csr_f32(const int32_t* rowptr, const int32_t* col, const float* val, const float* X, float* Y,
                               int32_t n_rows, int32_t D, md_stream_t stream) {
  if (!rowptr || !col || !X || !Y || n_rows < 0 || D <= 0 || (D & 3)) return MD_EINVAL;
  if (n_rows == 0) return 0;
  const int D4 = D / 4;
  if (D4 > 128) return MD_ETOOBIG;
  int dev = 0, nsm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  const int wpb = 8;
  const int rpb = D4 <= 16 ? 2 * wpb : wpb;  // rows per block
  long long want = ((long long)n_rows + rpb - 1) / rpb;
  int grid = (int)(want < (long long)nsm * 8 ? want : (long long)nsm * 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (D4 <= 16) md::spmm_csr_kernel<1, 16><<<grid, 256, 0, st>>>(rowptr, col, val, X, Y, n_rows, D4);
  else if (D4 <= 32) md::spmm_csr_kernel<1, 32><<<grid, 256, 0, st>>>(rowptr, col, val, X, Y, n_rows, D4);
  else if (D4 <= 64) md::spmm_csr_kernel<2, 32><<<grid, 256, 0, st>>>(rowptr, col, val, X, Y, n_rows, D4);
  else md::spmm_csr_kernel<4, 32><<<grid, 256, 0, st>>>(rowptr, col, val, X, Y, n_rows, D4);
  MD_CHECK_LAUNCH();
  return 0;
}
