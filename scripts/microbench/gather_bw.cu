// Microbenchmark: ceiling of random 256-byte row gathers on B200 when the table does not fit in L2 (C4: 256 MB per
// layer).  Two ways of keeping bytes in flight:
//   ldg   register loads, one row per half-warp instruction, G independent loads per lane before the adds
//   ring  cp.async (LDGSTS, 16 B per lane) into a per-warp shared-memory ring of R slots x 16 rows, summed from smem
// Indices are read coalesced (one per lane) and broadcast by shuffle, so they add no dependent latency per row.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ void cp16(void* smem, const void* g) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(a), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// every warp walks `batches` batches of 32 row ids; out gets one float4 per lane per batch
template <int G>
__global__ void __launch_bounds__(1024) ldg_kernel(const float4* __restrict__ H, const int* __restrict__ idx, float4* out,
                                                   int batches) {
  const int lane = threadIdx.x & 31, half = lane >> 4, l16 = lane & 15;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int* my = idx + (size_t)gw * batches * 32;
  float4 acc = make_float4(0, 0, 0, 0);
  for (int b = 0; b < batches; ++b) {
    const int id = my[b * 32 + lane];
#pragma unroll
    for (int j0 = 0; j0 < 16; j0 += G) {
      float4 v[G];
#pragma unroll
      for (int j = 0; j < G; ++j) {
        const int u = __shfl_sync(0xffffffffu, id, 16 * half + j0 + j);
        v[j] = __ldg(H + (size_t)u * 16 + l16);
      }
#pragma unroll
      for (int j = 0; j < G; ++j) acc.x += v[j].x, acc.y += v[j].y, acc.z += v[j].z, acc.w += v[j].w;
    }
  }
  out[(size_t)gw * 32 + lane] = acc;
}

// R slots of 32 rows (8 KB) per warp
template <int R>
__global__ void ring_kernel(const float4* __restrict__ H, const int* __restrict__ idx, float4* out, int batches) {
  extern __shared__ float4 ring[];
  const int lane = threadIdx.x & 31, half = lane >> 4, l16 = lane & 15, warp = threadIdx.x >> 5;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int* my = idx + (size_t)gw * batches * 32;
  float4* wr = ring + (size_t)warp * R * 512;  // 32 rows x 16 float4 per slot
  float4 acc = make_float4(0, 0, 0, 0);
  auto issue = [&](int b) {
    const int id = my[b * 32 + lane];
    float4* slot = wr + (b % R) * 512;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int u = __shfl_sync(0xffffffffu, id, 16 * half + j);
      cp16(slot + (2 * j + half) * 16 + l16, H + (size_t)u * 16 + l16);
    }
    cp_commit();
  };
  for (int b = 0; b < R - 1 && b < batches; ++b) issue(b);
  for (int b = 0; b < batches; ++b) {
    if (b + R - 1 < batches) issue(b + R - 1);
    else cp_commit();
    cp_wait<R - 1>();
    __syncwarp();
    const float4* slot = wr + (b % R) * 512;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float4 v = slot[(2 * j + half) * 16 + l16];
      acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
    }
    __syncwarp();
  }
  out[(size_t)gw * 32 + lane] = acc;
}

// cp.async.bulk: one 256-byte bulk copy per LANE (32 rows per warp instruction), completion on a per-slot mbarrier
__device__ __forceinline__ unsigned sm32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int R>
__global__ void bulk_kernel(const float4* __restrict__ H, const int* __restrict__ idx, float4* out, int batches) {
  extern __shared__ __align__(128) float4 ring[];
  __shared__ unsigned long long bars[32][R];
  const int lane = threadIdx.x & 31, half = lane >> 4, l16 = lane & 15, warp = threadIdx.x >> 5;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int* my = idx + (size_t)gw * batches * 32;
  float4* wr = ring + (size_t)warp * R * 512;
  if (lane == 0)
    for (int r = 0; r < R; ++r) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sm32(&bars[warp][r])));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  float4 acc = make_float4(0, 0, 0, 0);
  auto issue = [&](int b) {
    const int id = my[b * 32 + lane];
    float4* slot = wr + (b % R) * 512;
    const unsigned bar = sm32(&bars[warp][b % R]);
    if (lane == 0) asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(8192) : "memory");
    __syncwarp();
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sm32(slot + lane * 16)),
                 "l"(H + (size_t)id * 16), "r"(256), "r"(bar)
                 : "memory");
  };
  for (int b = 0; b < R - 1 && b < batches; ++b) issue(b);
  for (int b = 0; b < batches; ++b) {
    if (b + R - 1 < batches) issue(b + R - 1);
    const unsigned bar = sm32(&bars[warp][b % R]);
    const unsigned parity = (b / R) & 1;
    unsigned ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    const float4* slot = wr + (b % R) * 512;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float4 v = slot[(2 * j + half) * 16 + l16];
      acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
    }
    __syncwarp();
  }
  out[(size_t)gw * 32 + lane] = acc;
}

int main() {
  const int nsm = 148;
  for (int n_rows : {1 << 20, 3 << 20}) {
    float4* H;
    cudaMalloc(&H, (size_t)n_rows * 256);
    cudaMemset(H, 0, (size_t)n_rows * 256);
    const int total_rows = 24 << 20;  // gathers per run
    std::vector<int> h(total_rows);
    srand(1);
    for (auto& x : h) x = (int)(((unsigned)rand() * 32768u + (unsigned)rand()) % (unsigned)n_rows);
    int* idx;
    cudaMalloc(&idx, (size_t)total_rows * 4);
    cudaMemcpy(idx, h.data(), (size_t)total_rows * 4, cudaMemcpyHostToDevice);
    float4* out;
    cudaMalloc(&out, (size_t)nsm * 2048 * 16);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    auto report = [&](const char* tag, int warps_per_sm, float ms) {
      printf("H=%4d MB %-18s warps/SM=%2d: %7.1f us  %.2f TB/s of gathered rows\n", n_rows / 4096, tag, warps_per_sm, ms * 1e3,
             (double)total_rows * 256 / (ms * 1e-3) / 1e12);
    };
#define RUN_LDG(G, CTAS, THREADS)                                                     \
  {                                                                                   \
    const int warps = nsm * CTAS * THREADS / 32, batches = total_rows / 32 / warps;   \
    float ms = 0;                                                                     \
    for (int rep = 0; rep < 3; ++rep) {                                               \
      cudaEventRecord(e0);                                                            \
      ldg_kernel<G><<<nsm * CTAS, THREADS>>>(H, idx, out, batches);                   \
      cudaEventRecord(e1);                                                            \
      cudaEventSynchronize(e1);                                                       \
      cudaEventElapsedTime(&ms, e0, e1);                                              \
    }                                                                                 \
    report("ldg G=" #G, CTAS* THREADS / 32, ms * (float)total_rows / (float)(batches * 32 * warps)); \
  }
    RUN_LDG(4, 1, 512)
    RUN_LDG(8, 1, 512)
    RUN_LDG(16, 1, 512)
    RUN_LDG(4, 1, 1024)
    RUN_LDG(8, 1, 1024)
    RUN_LDG(16, 1, 1024)
    RUN_LDG(4, 2, 1024)
    RUN_LDG(8, 2, 1024)
    RUN_LDG(16, 2, 1024)
#define RUN_RING(R, THREADS)                                                                          \
  {                                                                                                   \
    const int warps = nsm * THREADS / 32, batches = total_rows / 32 / warps;                          \
    const int smem = THREADS / 32 * R * 8192;                                                         \
    cudaFuncSetAttribute(ring_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);          \
    float ms = 0;                                                                                     \
    for (int rep = 0; rep < 3; ++rep) {                                                               \
      cudaEventRecord(e0);                                                                            \
      ring_kernel<R><<<nsm, THREADS, smem>>>(H, idx, out, batches);                                   \
      cudaEventRecord(e1);                                                                            \
      cudaEventSynchronize(e1);                                                                       \
      cudaEventElapsedTime(&ms, e0, e1);                                                              \
    }                                                                                                 \
    char tag[64];                                                                                     \
    snprintf(tag, 64, "ring R=%d (%d KB)", R, smem / 1024);                                           \
    report(tag, THREADS / 32, ms * (float)total_rows / (float)(batches * 32 * warps));                \
  }
    RUN_RING(2, 256)
    RUN_RING(3, 256)
    RUN_RING(4, 128)
    RUN_RING(6, 128)
    RUN_RING(12, 64)
    RUN_RING(24, 32)
#define RUN_BULK(R, THREADS)                                                                          \
  {                                                                                                   \
    const int warps = nsm * THREADS / 32, batches = total_rows / 32 / warps;                          \
    const int smem = THREADS / 32 * R * 8192;                                                         \
    cudaFuncSetAttribute(bulk_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);          \
    float ms = 0;                                                                                     \
    for (int rep = 0; rep < 3; ++rep) {                                                               \
      cudaEventRecord(e0);                                                                            \
      bulk_kernel<R><<<nsm, THREADS, smem>>>(H, idx, out, batches);                                   \
      cudaEventRecord(e1);                                                                            \
      cudaEventSynchronize(e1);                                                                       \
      cudaEventElapsedTime(&ms, e0, e1);                                                              \
    }                                                                                                 \
    char tag[64];                                                                                     \
    snprintf(tag, 64, "bulk R=%d (%d KB)", R, smem / 1024);                                           \
    report(tag, THREADS / 32, ms * (float)total_rows / (float)(batches * 32 * warps));                \
  }
    RUN_BULK(2, 256)
    RUN_BULK(3, 256)
    RUN_BULK(4, 128)
    RUN_BULK(6, 128)
    RUN_BULK(12, 64)
    RUN_BULK(24, 32)
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) printf("error: %s\n", cudaGetErrorString(e));
    cudaFree(H), cudaFree(idx), cudaFree(out);
  }
  return 0;
}
