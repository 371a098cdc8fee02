// CUDA-core implicit-GEMM convolution (fp32 accumulate).  This is the FP32CHECK arithmetic of every
// conv / transposed conv (north_star's "<=1e-4 check mode"); the h16 product path uses
// conv_tc.cuh instead.
//
// Semantics follow Keras Conv2D / Conv2DTranspose with padding='same' as used by
// /root/reference/model.py:119,153 :
//   conv  : out[oy,ox,o] = sum_{kh,kw,c} in[oy*s + kh - pad_t, ox*s + kw - pad_l, c] * W[kh,kw,c,o]
//           with TF SAME padding (pad_before = total/2), zero outside;
//   convT : out[Y,X,o]  = sum over (kh,kw) with (Y-kh),(X-kw) even and >=0 and in range of
//           in[(Y-kh)/2,(X-kw)/2,c] * W[kh,kw,c,o]   (3x3, stride 2, output = 2x input).
// Weights are canonical fp32 [tap][cin][cout] (BN folded by the converter).
#pragma once
#include "common.cuh"

namespace bsr {

struct ConvDirectParams {
  const void* in;     // [N, H, W, in_ld] (channels in_coff .. in_coff+cin)
  int in_ld, in_coff, cin;
  int H, W, OH, OW, N;
  int kh, kw, stride, pad_t, pad_l;
  int transposed;
  const float* w;     // [kh*kw][cin][cout]
  int cout;
};

constexpr int CD_TM = 64, CD_TN = 64, CD_TK = 16;

template <typename TIn, typename T>
__global__ void __launch_bounds__(256) conv_direct_kernel(ConvDirectParams p, EpiParams e) {
  __shared__ float As[CD_TK][CD_TM + 4];
  __shared__ float Bs[CD_TK][CD_TN];
  const int t = threadIdx.x;
  const long long M = (long long)p.N * p.OH * p.OW;
  const long long m0 = (long long)blockIdx.x * CD_TM;
  const int co0 = blockIdx.y * CD_TN;

  // A-load role: pixel a_m (0..63), channel group a_k (0..3) -> 4 channels
  const int a_m = t >> 2, a_k = (t & 3) * 4;
  long long am = m0 + a_m;
  const bool a_ok = am < M;
  int an = 0, aoy = 0, aox = 0;
  if (a_ok) {
    an = (int)(am / ((long long)p.OH * p.OW));
    int r = (int)(am % ((long long)p.OH * p.OW));
    aoy = r / p.OW;
    aox = r % p.OW;
  }
  // B-load role: k row b_k (0..15), 4 couts at b_n
  const int b_k = t >> 4, b_n = (t & 15) * 4;
  const int tm = (t >> 4) * 4, tn = (t & 15) * 4;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const TIn* in = (const TIn*)p.in;
  for (int tap = 0; tap < p.kh * p.kw; ++tap) {
    const int kh = tap / p.kw, kw = tap % p.kw;
    int iy, ix;
    bool valid = a_ok;
    if (!p.transposed) {
      iy = aoy * p.stride + kh - p.pad_t;
      ix = aox * p.stride + kw - p.pad_l;
    } else {
      int ty = aoy - kh, tx = aox - kw;
      valid = valid && ty >= 0 && tx >= 0 && ((ty | tx) & 1) == 0;
      iy = ty >> 1;
      ix = tx >> 1;
    }
    valid = valid && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
    const size_t a_base = valid ? (((size_t)an * p.H + iy) * p.W + ix) * p.in_ld + p.in_coff : 0;
    const float* wt = p.w + (size_t)tap * p.cin * p.cout;
    for (int c0 = 0; c0 < p.cin; c0 += CD_TK) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int c = c0 + a_k + i;
        As[a_k + i][a_m] = (valid && c < p.cin) ? ldf<TIn>(in, a_base + c) : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int c = c0 + b_k, co = co0 + b_n + j;
        Bs[b_k][b_n + j] = (c < p.cin && co < p.cout) ? __ldg(wt + (size_t)c * p.cout + co) : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < CD_TK; ++k) {
        float a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = As[k][tm + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = Bs[k][tn + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    long long m = m0 + tm + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) epi_store<T>(e, (size_t)m, co0 + tn + j, acc[i][j]);
  }
}

}  // namespace bsr
