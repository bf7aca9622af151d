// CSR SpMM, the operator-level boundary: drop-in for torch_sparse.spmm (torch-sparse 0.6.18) as the reference
// calls it at MultiDismantler_net_graphsage.py:139,146,199,290,297,350,377 and MultiDismantler_torch.py:420.
//   Y[r, :] = sum_{e in row r} val[e] * X[col[e], :]        (val == NULL: all ones, aggregatorID = 0)
// One warp per row; each lane owns D/4 <= ... float4 column chunks; neighbours are taken in CSR order so the
// f32 sum is deterministic and matches a sequential scatter-add in COO order.
#include "md_common.cuh"

namespace md {

template <int VEC>  // number of float4 chunks per lane (D = 128 * VEC max)
__global__ void __launch_bounds__(256) spmm_csr_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                      const float* __restrict__ val, const float* __restrict__ X,
                                                      float* __restrict__ Y, int n_rows, int D4) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < n_rows; row += gridDim.x * wpb) {
    const int b = rowptr[row], e = rowptr[row + 1];
    float4 acc[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int c0 = b; c0 < e; c0 += 32) {
      int u = 0;
      float w = 1.f;
      if (c0 + lane < e) {
        u = col[c0 + lane];
        if (val) w = val[c0 + lane];
      }
      const int cnt = min(32, e - c0);
      for (int j0 = 0; j0 < cnt; j0 += 4) {
        float4 x[4][VEC];
        float ws[4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int uu = __shfl_sync(0xffffffffu, u, (j0 + jj) & 31);
          ws[jj] = __shfl_sync(0xffffffffu, w, (j0 + jj) & 31);
          const bool ok = j0 + jj < cnt;
#pragma unroll
          for (int j = 0; j < VEC; ++j) {
            x[jj][j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ok && lane + 32 * j < D4) x[jj][j] = __ldg(reinterpret_cast<const float4*>(X) + (size_t)uu * D4 + lane + 32 * j);
          }
        }
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
          if (j0 + jj < cnt) {
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
#pragma unroll
    for (int j = 0; j < VEC; ++j)
      if (lane + 32 * j < D4) reinterpret_cast<float4*>(Y)[(size_t)row * D4 + lane + 32 * j] = acc[j];
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
  long long want = ((long long)n_rows + wpb - 1) / wpb;
  int grid = (int)(want < (long long)nsm * 8 ? want : (long long)nsm * 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (D4 <= 32) md::spmm_csr_kernel<1><<<grid, 256, 0, st>>>(rowptr, col, val, X, Y, n_rows, D4);
  else if (D4 <= 64) md::spmm_csr_kernel<2><<<grid, 256, 0, st>>>(rowptr, col, val, X, Y, n_rows, D4);
  else md::spmm_csr_kernel<4><<<grid, 256, 0, st>>>(rowptr, col, val, X, Y, n_rows, D4);
  MD_CHECK_LAUNCH();
  return 0;
}
