// Non-local attention on CUDA cores, fp32 math (FP32CHECK mode): y = softmax(theta . phi^T) . g,
// logits NOT scaled (/root/reference/model.py:51-53).  One CTA per (image, 32 queries): the full
// 32x1024 logit strip lives in shared memory, so the [1024,1024] matrix never reaches HBM.
// Layouts (shared with the tensor-core kernel): QK[n][s][256] = theta(0:128) | phi(128:256);
// VT[n][128][S] = g transposed; O[n][s][o_ld] channels 0:128.
#pragma once
#include "common.cuh"

namespace bsr {

constexpr int AS_Q = 32;       // queries per CTA
constexpr int AS_S = 1024;     // tokens (32x32)
constexpr int AS_D = 128;

template <typename T>
__global__ void __launch_bounds__(256) attention_simple_kernel(const T* __restrict__ qk, const T* __restrict__ vt,
                                                               T* __restrict__ o, int o_ld) {
  extern __shared__ float sm[];
  float* sc = sm;                            // [32][1024]
  float* qs = sc + AS_Q * AS_S;              // [32][128]
  float* vs = qs + AS_Q * AS_D;              // [128][33]
  const int t = threadIdx.x;
  const int n = blockIdx.y, q0 = blockIdx.x * AS_Q;
  const T* qk_n = qk + (size_t)n * AS_S * 256;
  const T* vt_n = vt + (size_t)n * AS_D * AS_S;

  for (int i = t; i < AS_Q * AS_D; i += 256) qs[i] = ldf<T>(qk_n, (size_t)(q0 + i / AS_D) * 256 + (i % AS_D));
  __syncthreads();

  // logits: each thread owns keys t, t+256, ...
  for (int k = t; k < AS_S; k += 256) {
    float acc[AS_Q];
#pragma unroll
    for (int q = 0; q < AS_Q; ++q) acc[q] = 0.f;
    const T* kr = qk_n + (size_t)k * 256 + 128;
    for (int d = 0; d < AS_D; d += 4) {
      float k0 = ldf<T>(kr, d), k1 = ldf<T>(kr, d + 1), k2 = ldf<T>(kr, d + 2), k3 = ldf<T>(kr, d + 3);
#pragma unroll
      for (int q = 0; q < AS_Q; ++q) {
        const float* qq = qs + q * AS_D + d;
        acc[q] = fmaf(qq[0], k0, acc[q]);
        acc[q] = fmaf(qq[1], k1, acc[q]);
        acc[q] = fmaf(qq[2], k2, acc[q]);
        acc[q] = fmaf(qq[3], k3, acc[q]);
      }
    }
#pragma unroll
    for (int q = 0; q < AS_Q; ++q) sc[q * AS_S + k] = acc[q];
  }
  __syncthreads();

  // softmax per row: warp w handles rows w*4 .. w*4+3
  const int warp = t >> 5, lane = t & 31;
  for (int r = warp * 4; r < warp * 4 + 4; ++r) {
    float* row = sc + r * AS_S;
    float mx = -INFINITY;
    for (int k = lane; k < AS_S; k += 32) mx = fmaxf(mx, row[k]);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, s));
    float sum = 0.f;
    for (int k = lane; k < AS_S; k += 32) {
      float ev = expf(row[k] - mx);
      row[k] = ev;
      sum += ev;
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, s);
    float inv = 1.f / sum;
    for (int k = lane; k < AS_S; k += 32) row[k] *= inv;
  }
  __syncthreads();

  // O = P . V : thread -> channel d = t%128, query half (t/128) of 16 queries
  const int d = t & 127, qh = (t >> 7) * 16;
  float out[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) out[q] = 0.f;
  for (int k0 = 0; k0 < AS_S; k0 += 32) {
    for (int i = t; i < AS_D * 32; i += 256) {
      int dd = i >> 5, kk = i & 31;
      vs[dd * 33 + kk] = ldf<T>(vt_n, (size_t)dd * AS_S + k0 + kk);
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < 32; ++kk) {
      float v = vs[d * 33 + kk];
#pragma unroll
      for (int q = 0; q < 16; ++q) out[q] = fmaf(sc[(qh + q) * AS_S + k0 + kk], v, out[q]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int q = 0; q < 16; ++q)
    stf<T>(o, ((size_t)n * AS_S + q0 + qh + q) * o_ld + d, out[q]);
}

constexpr size_t kAttnSimpleSmem = (size_t)(AS_Q * AS_S + AS_Q * AS_D + AS_D * 33) * sizeof(float);

}  // namespace bsr
