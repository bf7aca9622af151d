// Shared device helpers for the sm_100a rollout kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "mdcommunity_b200.h"

#define MD_CHECK_LAUNCH()                         \
  do {                                            \
    cudaError_t e__ = cudaGetLastError();         \
    if (e__ != cudaSuccess) return (int)e__;      \
  } while (0)

namespace md {

constexpr int D = MD_EMBED;
constexpr int TL_N = 64;    // md_rollout_tail: most active nodes a graph may have at hand-over
constexpr int TL_E = 4096;  //                  most alive directed entries per layer at hand-over
constexpr int FLAG_LEFT = 8;  // md_state_t.flags slot: graphs that left the lock-step iteration
constexpr int TILE_M = MD_TILE_M;
// row-major A tile [TILE_M][2*D] in shared memory, padded so that the 8 rows a thread reads
// (ty + 16*i) and the 4 row groups of a warp fall into different banks for 16-byte accesses
constexpr int A_LD = 2 * D + 4;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_max_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// sum over the 8 consecutive lanes that share a row in the 16x8 thread layout of the tile GEMMs
__device__ __forceinline__ float oct_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}

// monotone float -> uint map (larger float => larger uint), used for the packed arg-max key
__device__ __forceinline__ uint32_t orderable(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float from_orderable(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// ---- host side: caches keyed by the current device --------------------------------------------------------------
// cudaFuncAttributeMaxDynamicSharedMemorySize is per device and SM counts may differ between devices, so the
// "configured once" flags of the launchers are kept per device (a second GPU in the same process gets its own).
constexpr int MAX_DEVICES = 64;
inline int cur_device() {
  int d = 0;
  cudaGetDevice(&d);
  return (d < 0 || d >= MAX_DEVICES) ? 0 : d;
}
template <class T>
struct PerDevice {
  T v[MAX_DEVICES] = {};
  T& get() { return v[cur_device()]; }
};

}  // namespace md
