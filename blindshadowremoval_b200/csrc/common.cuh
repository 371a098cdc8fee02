// Shared device helpers: activation storage types, the fused conv epilogue, small math.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace bsr {

constexpr float kLeaky = 0.3f;       // keras LeakyReLU default alpha (model.py:90-92,130,161)
constexpr float kHoleThr = 0.1f;     // model.py:256
constexpr float kGrayR = 0.2989f, kGrayG = 0.5870f, kGrayB = 0.1140f;   // tf.image.rgb_to_grayscale

typedef __nv_bfloat16 bf16;

template <typename T> __device__ __forceinline__ float ldf(const T* p, size_t i);
template <> __device__ __forceinline__ float ldf<float>(const float* p, size_t i) { return p[i]; }
template <> __device__ __forceinline__ float ldf<bf16>(const bf16* p, size_t i) { return __bfloat162float(p[i]); }
template <typename T> __device__ __forceinline__ void stf(T* p, size_t i, float v);
template <> __device__ __forceinline__ void stf<float>(float* p, size_t i, float v) { p[i] = v; }
template <> __device__ __forceinline__ void stf<bf16>(bf16* p, size_t i, float v) { p[i] = __float2bfloat16_rn(v); }

__device__ __forceinline__ float leaky(float v) { return v >= 0.f ? v : kLeaky * v; }

enum OutMode : int {
  OUT_T = 0,     // activation type T, NHWC at out[pix*out_ld + out_coff + c]
  OUT_F32 = 1,   // float, same addressing
  OUT_QKV = 2    // non-local projections: c<256 -> QK[pix*256 + c] (theta|phi), c>=256 -> V^T[n][c-256][pix%S]
};

// Fused epilogue of every convolution (BN already folded into W / bias):
//   v = acc + bias[c]            (c < cout, else 0)
//   v += res1[pix, c] (c<res1_c) ; v += res2[pix, c] (c<res2_c)      residual / skip adds with
//                                                                     channel zero-extension
//   v = leaky(v) if act
//   store for c < out_c
// `pix` is the linear output pixel index (n*OH + oy)*OW + ox.
struct EpiParams {
  const float* bias;
  int cout;
  int act;
  const void* res1; int res1_ld; int res1_c;
  const void* res2; int res2_ld; int res2_c;
  void* out; int out_ld; int out_coff; int out_c;
  int out_mode;
  void* out2;        // OUT_QKV: V^T base
  int spatial;       // OUT_QKV: pixels per image (1024)
};

template <typename T>
__device__ __forceinline__ float epi_value(const EpiParams& e, size_t pix, int c, float acc) {
  float v = (c < e.cout) ? acc + __ldg(e.bias + c) : 0.f;
  if (e.res1 != nullptr && c < e.res1_c) v += ldf<T>((const T*)e.res1, pix * e.res1_ld + c);
  if (e.res2 != nullptr && c < e.res2_c) v += ldf<T>((const T*)e.res2, pix * e.res2_ld + c);
  if (e.act) v = leaky(v);
  return v;
}

template <typename T>
__device__ __forceinline__ void epi_store(const EpiParams& e, size_t pix, int c, float acc) {
  if (c >= e.out_c) return;
  float v = epi_value<T>(e, pix, c, acc);
  if (e.out_mode == OUT_F32) {
    ((float*)e.out)[pix * e.out_ld + e.out_coff + c] = v;
  } else if (e.out_mode == OUT_T) {
    stf<T>((T*)e.out, pix * e.out_ld + e.out_coff + c, v);
  } else {
    if (c < 256) {
      stf<T>((T*)e.out, pix * 256 + c, v);
    } else {
      size_t n = pix / e.spatial, s = pix % e.spatial;
      stf<T>((T*)e.out2, (n * 128 + (c - 256)) * e.spatial + s, v);
    }
  }
}

}  // namespace bsr
