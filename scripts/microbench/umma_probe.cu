// Probe for the tcgen05 building blocks of the fused embedding kernel (run on a B200):
//   D[128x64] = [P | H] (128 x 64, fp32) @ [W13c ; W23c] (64 x 64, fp32) through kind::tf32 with the hi/lo split
//   (A raw fp32 in SWIZZLE_128B K-major tiles, B = hi and lo planes) plus one kind::f16 (bf16) correction pass for the
//   A remainders, accumulators in TMEM, read back with tcgen05.ld.
// Checks (a) descriptors / layouts (result matches at all), (b) whether tf32 operands are truncated or rounded,
// (c) the accuracy of the 3-term product against an f64 reference, (d) MMA issue->commit latency.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I mdcommunity_b200/csrc -I include -o umma_probe umma_probe.cu
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "md_umma.cuh"

using namespace md::umma;

struct Smem {
  alignas(1024) unsigned char P[16384];
  alignas(1024) unsigned char H[16384];
  alignas(1024) unsigned char LO[16384];
  alignas(1024) unsigned char B13hi[8192];
  alignas(1024) unsigned char B13lo[8192];
  alignas(1024) unsigned char B23hi[8192];
  alignas(1024) unsigned char B23lo[8192];
  alignas(1024) unsigned char Bbf[8192];
  uint64_t bar;
  uint32_t tmem;
};

__global__ void __launch_bounds__(160, 1) probe(const float* P, const float* H, const float* W13, const float* W23, float* Dout,
                                                int mode, long long* cycles) {
  extern __shared__ __align__(1024) unsigned char raw[];
  Smem& sm = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(&sm.bar, 1);
    mbar_fence_init();
  }
  if (warp == 4) tmem_alloc(&sm.tmem, 64);
  // fill the A tiles: thread r < 128 owns row r
  if (tid < 128) {
    const int r = tid;
    for (int ch = 0; ch < 8; ++ch) {
      float4 p = *reinterpret_cast<const float4*>(P + r * 32 + 4 * ch);
      float4 h = *reinterpret_cast<const float4*>(H + r * 32 + 4 * ch);
      *reinterpret_cast<float4*>(sm.P + sw128_off(r, ch)) = p;
      *reinterpret_cast<float4*>(sm.H + sw128_off(r, ch)) = h;
      // remainders as bf16: LO row = [P_lo (32) | H_lo (32)], 8 bytes per 4 values
      float pl[4] = {p.x - tf32_hi(p.x), p.y - tf32_hi(p.y), p.z - tf32_hi(p.z), p.w - tf32_hi(p.w)};
      float hl[4] = {h.x - tf32_hi(h.x), h.y - tf32_hi(h.y), h.z - tf32_hi(h.z), h.w - tf32_hi(h.w)};
      uint2 pa, hb;  // 4 values = 8 bytes: chunk (16 B) index ch/2, half (ch&1)
      if (mode == 2) {
        __half2 a0 = __floats2half2_rn(pl[0], pl[1]), a1 = __floats2half2_rn(pl[2], pl[3]);
        __half2 b0 = __floats2half2_rn(hl[0], hl[1]), b1 = __floats2half2_rn(hl[2], hl[3]);
        pa = make_uint2(*reinterpret_cast<uint32_t*>(&a0), *reinterpret_cast<uint32_t*>(&a1));
        hb = make_uint2(*reinterpret_cast<uint32_t*>(&b0), *reinterpret_cast<uint32_t*>(&b1));
      } else {
        __nv_bfloat162 a0 = __floats2bfloat162_rn(pl[0], pl[1]), a1 = __floats2bfloat162_rn(pl[2], pl[3]);
        __nv_bfloat162 b0 = __floats2bfloat162_rn(hl[0], hl[1]), b1 = __floats2bfloat162_rn(hl[2], hl[3]);
        pa = make_uint2(*reinterpret_cast<uint32_t*>(&a0), *reinterpret_cast<uint32_t*>(&a1));
        hb = make_uint2(*reinterpret_cast<uint32_t*>(&b0), *reinterpret_cast<uint32_t*>(&b1));
      }
      *reinterpret_cast<uint2*>(sm.LO + sw128_off(r, ch >> 1) + 8 * (ch & 1)) = pa;
      *reinterpret_cast<uint2*>(sm.LO + sw128_off(r, 4 + (ch >> 1)) + 8 * (ch & 1)) = hb;
    }
    // B tiles: row n < 64 = output column, K contiguous.  W is [k][n] row-major in global memory.
    if (r < 64) {
      const int n = r;
      for (int k = 0; k < 32; ++k) {
        const float w13 = W13[k * 64 + n], w23 = W23[k * 64 + n];
        const uint32_t off = sw128_off(n, k >> 2) + 4 * (k & 3);
        *reinterpret_cast<float*>(sm.B13hi + off) = tf32_hi(w13);
        *reinterpret_cast<float*>(sm.B13lo + off) = w13 - tf32_hi(w13);
        *reinterpret_cast<float*>(sm.B23hi + off) = tf32_hi(w23);
        *reinterpret_cast<float*>(sm.B23lo + off) = w23 - tf32_hi(w23);
        // bf16 plane [n][64]: k < 32 -> W13, k >= 32 -> W23
        if (mode == 2) {
          *reinterpret_cast<__half*>(sm.Bbf + sw128_off(n, (2 * k) >> 4) + ((2 * k) & 15)) = __float2half_rn(w13);
          *reinterpret_cast<__half*>(sm.Bbf + sw128_off(n, (2 * (k + 32)) >> 4) + ((2 * (k + 32)) & 15)) = __float2half_rn(w23);
        } else {
          *reinterpret_cast<__nv_bfloat16*>(sm.Bbf + sw128_off(n, (2 * k) >> 4) + ((2 * k) & 15)) = __float2bfloat16_rn(w13);
          *reinterpret_cast<__nv_bfloat16*>(sm.Bbf + sw128_off(n, (2 * (k + 32)) >> 4) + ((2 * (k + 32)) & 15)) = __float2bfloat16_rn(w23);
        }
      }
    }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = sm.tmem;
  if (warp == 4) {
    if (lane == 0) {
      const long long t0 = clock64();
      const uint32_t idt = idesc(FMT_TF32, 128, 64), idb = idesc(mode == 2 ? FMT_F16 : FMT_BF16, 128, 64);
      uint32_t acc = 0;
      for (int k = 0; k < 4; ++k) {
        mma_tf32(tbase, desc_k_sw128(smem_u32(sm.P), 32 * k), desc_k_sw128(smem_u32(sm.B13hi), 32 * k), idt, acc);
        acc = 1;
      }
      for (int k = 0; k < 4; ++k) mma_tf32(tbase, desc_k_sw128(smem_u32(sm.H), 32 * k), desc_k_sw128(smem_u32(sm.B23hi), 32 * k), idt, 1);
      if (mode >= 1) {
        for (int k = 0; k < 4; ++k) mma_tf32(tbase, desc_k_sw128(smem_u32(sm.P), 32 * k), desc_k_sw128(smem_u32(sm.B13lo), 32 * k), idt, 1);
        for (int k = 0; k < 4; ++k) mma_tf32(tbase, desc_k_sw128(smem_u32(sm.H), 32 * k), desc_k_sw128(smem_u32(sm.B23lo), 32 * k), idt, 1);
        for (int k = 0; k < 4; ++k) mma_f16(tbase, desc_k_sw128(smem_u32(sm.LO), 32 * k), desc_k_sw128(smem_u32(sm.Bbf), 32 * k), idb, 1);
      }
      const long long t1 = clock64();
      mma_commit(&sm.bar);
      mbar_wait(&sm.bar, 0);
      const long long t2 = clock64();
      cycles[0] = t1 - t0;
      cycles[1] = t2 - t0;
    }
    __syncwarp();
  }
  if (warp < 4) {
    mbar_wait(&sm.bar, 0);
    tc_fence_after();
    const int row = 32 * warp + lane;
    for (int c0 = 0; c0 < 64; c0 += 16) {
      float v[16];
      tmem_ld16(tbase + ((uint32_t)(32 * warp) << 16) + c0, v);
      tmem_ld_wait();
      for (int i = 0; i < 16; ++i) Dout[row * 64 + c0 + i] = v[i];
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) tmem_dealloc(tbase, 64);
}

static float trunc_tf32(float x) { return tf32_hi(x); }
static float rna_tf32(float x) {
  union { float f; uint32_t u; } c;
  c.f = x;
  c.u += 0x1000u;
  c.u &= 0xFFFFE000u;
  return c.f;
}

int main() {
  const int M = 128;
  std::vector<float> P(M * 32), H(M * 32), W13(32 * 64), W23(32 * 64), D(M * 64);
  srand(1);
  auto rnd = [] { return (float)rand() / RAND_MAX; };
  for (auto& x : P) x = 6.f * rnd();
  for (auto& x : H) x = rnd();
  for (auto& x : W13) x = 2.f * rnd() - 1.f;
  for (auto& x : W23) x = 2.f * rnd() - 1.f;
  float *dP, *dH, *dW13, *dW23, *dD;
  long long* dC;
  cudaMalloc(&dP, P.size() * 4);
  cudaMalloc(&dH, H.size() * 4);
  cudaMalloc(&dW13, W13.size() * 4);
  cudaMalloc(&dW23, W23.size() * 4);
  cudaMalloc(&dD, D.size() * 4);
  cudaMalloc(&dC, 16);
  cudaMemcpy(dP, P.data(), P.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dH, H.data(), H.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dW13, W13.data(), W13.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dW23, W23.data(), W23.size() * 4, cudaMemcpyHostToDevice);
  const int smem = sizeof(Smem) + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int mode = 0; mode < 3; ++mode) {  // 0: tf32 hi x hi only, 1: 3-term with bf16 remainders, 2: 3-term with fp16 remainders
    cudaMemset(dD, 0, D.size() * 4);
    probe<<<1, 160, smem>>>(dP, dH, dW13, dW23, dD, mode, dC);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("mode %d: CUDA error %s\n", mode, cudaGetErrorString(e));
      return 1;
    }
    long long cyc[2];
    cudaMemcpy(cyc, dC, 16, cudaMemcpyDeviceToHost);
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double e_full = 0, e_trunc = 0, e_rna = 0, ref_max = 0;
    for (int r = 0; r < M; ++r)
      for (int n = 0; n < 64; ++n) {
        double full = 0, tr = 0, rn = 0;
        for (int k = 0; k < 32; ++k) {
          full += (double)P[r * 32 + k] * W13[k * 64 + n] + (double)H[r * 32 + k] * W23[k * 64 + n];
          tr += (double)trunc_tf32(P[r * 32 + k]) * trunc_tf32(W13[k * 64 + n]) + (double)trunc_tf32(H[r * 32 + k]) * trunc_tf32(W23[k * 64 + n]);
          rn += (double)rna_tf32(P[r * 32 + k]) * trunc_tf32(W13[k * 64 + n]) + (double)rna_tf32(H[r * 32 + k]) * trunc_tf32(W23[k * 64 + n]);
        }
        const double d = D[r * 64 + n];
        e_full = fmax(e_full, fabs(d - full));
        e_trunc = fmax(e_trunc, fabs(d - tr));
        e_rna = fmax(e_rna, fabs(d - rn));
        ref_max = fmax(ref_max, fabs(full));
      }
    printf("mode %d: max|D-f64| = %.3e  max|D-trunc_model| = %.3e  max|D-rna_model| = %.3e  (max|ref| %.2f)  issue %lld cyc, issue+complete %lld cyc\n",
           mode, e_full, e_trunc, e_rna, ref_max, cyc[0], cyc[1]);
  }
  return 0;
}
