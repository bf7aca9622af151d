// Microbenchmark: latency / throughput of 256-byte row gathers (one LDG.128 per half-warp) on B200.
// Each half-warp performs `rounds` dependent rounds; a round issues G independent row loads and sums them.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

template <int G, bool NC>
__global__ void gather_kernel(const float4* __restrict__ H, const int* __restrict__ idx, float4* out, int rounds,
                              int rows_per_cta_window, int n_rows) {
  const int half = (threadIdx.x & 31) >> 4, l16 = threadIdx.x & 15;
  const int hw = (threadIdx.x >> 5) * 2 + half;
  const int nhw = blockDim.x / 16;
  float4 acc = make_float4(0, 0, 0, 0);
  const int* my = idx + ((size_t)blockIdx.x * nhw + hw) * rounds * G;
  const size_t base = ((size_t)blockIdx.x * rows_per_cta_window) % n_rows;
  for (int r = 0; r < rounds; ++r) {
    float4 v[G];
#pragma unroll
    for (int j = 0; j < G; ++j) {
      const size_t row = (base + my[r * G + j]) % n_rows;
      if (NC) v[j] = __ldg(H + row * 16 + l16);
      else v[j] = H[row * 16 + l16];
    }
#pragma unroll
    for (int j = 0; j < G; ++j) { acc.x += v[j].x; acc.y += v[j].y; acc.z += v[j].z; acc.w += v[j].w; }
    // make the next round depend on this one (like a barrier / reduction would)
    if (acc.x == 123456.f) my += 1;
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int G>
void run(const float4* H, const int* idx, float4* out, int warps, int rounds, int window, int n_rows, const char* tag) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    gather_kernel<G, true><<<148, warps * 32>>>(H, idx, out, rounds, window, n_rows);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
  }
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double bytes = 148.0 * warps * 2 * rounds * G * 256;
  printf("%s warps=%2d G=%2d rounds=%d window=%d: %.1f us, %.0f ns/round, %.2f TB/s\n", tag, warps, G, rounds, window,
         ms * 1e3, ms * 1e6 / rounds, bytes / (ms * 1e-3) / 1e12);
}

int main() {
  const int n_rows = 1 << 20;  // 256 MB of rows
  float4* H; cudaMalloc(&H, (size_t)n_rows * 256); cudaMemset(H, 0, (size_t)n_rows * 256);
  const int max_idx = 148 * 64 * 64 * 16;
  std::vector<int> h(max_idx);
  float4* out; cudaMalloc(&out, 148 * 1024 * sizeof(float4));
  int* idx; cudaMalloc(&idx, max_idx * sizeof(int));
  for (int window : {1024, 1 << 20}) {
    srand(1);
    for (auto& x : h) x = rand() % window;
    cudaMemcpy(idx, h.data(), max_idx * sizeof(int), cudaMemcpyHostToDevice);
    const char* tag = window == 1024 ? "win256KB" : "win256MB";
    for (int warps : {1, 4, 16, 32}) {
      run<1>(H, idx, out, warps, 64, window, n_rows, tag);
      run<4>(H, idx, out, warps, 64, window, n_rows, tag);
      run<8>(H, idx, out, warps, 64, window, n_rows, tag);
      run<16>(H, idx, out, warps, 32, window, n_rows, tag);
    }
  }
  return 0;
}
