// Post-processing of FSRNet.test_step (/root/reference/train_test_GSC.py:436-725), the block that follows the generator
// call in the UCB evaluation (BASELINE config 2), as device kernels: resize to the crop size + pad (:438-476), region
// heuristics on the predicted shadow mask (:479-580), 4-connected component filter (:590-611, cv2.connectedComponents
// WithStats in the reference), nose rule (:650-662), composite + clip (:711-718), SSIM / PSNR vs ground truth (:724-725).
// No host round trip: every data-dependent decision of the reference (bounding boxes, area fractions, mean intensities)
// is a small statistics record per image that later kernels read; integer / fixed-point atomics keep the sums exact and
// order-independent, so results are bit-reproducible.
// Region masks are single-channel {0,1} bytes (the reference reads grey PNGs as three identical channels).
#pragma once
#include <limits.h>

#include "common.cuh"

namespace bsr {

constexpr int PP_IMG = 256, PP_PIX = PP_IMG * PP_IMG;
enum PpMask : int { PP_FACE_HAIR = 0, PP_FACE = 1, PP_MOUTH = 2, PP_NOSE = 3, PP_EYEBROW = 4, PP_EYE = 5, PP_GLASSES = 6, PP_NMASK = 7 };
// statistics record (32-bit slots; 64-bit fixed-point sums take two)
enum PpStat : int {
  PS_NOSE = 0, PS_MOUTH = 4, PS_BROW = 8, PS_FACE = 12, PS_FOREHEAD = 16,    // bounding boxes: rmin rmax cmin cmax
  PS_BROW_SUM = 20, PS_ROI_SUM = 21, PS_ROI_SHADOWED = 22, PS_CC_MAX = 23, PS_IMG2_SUM = 24, PS_NOSE_SUM = 25,
  PS_NOSE_SHADOW = 26, PS_ROI0_SHADOWED = 27,
  PS_MAB_FX = 32, PS_SHADOW_FX = 34, PS_SQERR_FX = 36, PS_SSIM_FX = 38, PS_WORDS = 64
};
constexpr double kPpFx = 1099511627776.0;      // 2^40 fixed point

struct PpPlanes {          // per-image scratch, all [256][256]
  float *tmp, *gt, *pred;  // [3] planes each, interleaved HWC like the inputs
  float *mp, *inten;       // mask_pred (after suppression), mean input intensity
  unsigned char *masks;    // [7][256][256] resized + rounded + padded
  unsigned char *detected, *img2;
  int *label, *csize, *chair;
  int* stats;              // [PS_WORDS]
};

__device__ __forceinline__ void pp_fx_add(int* stats, int slot, double v) {
  atomicAdd(reinterpret_cast<unsigned long long*>(stats + slot), (unsigned long long)(long long)llrint(v * kPpFx));
}
__device__ __forceinline__ double pp_fx_get(const int* stats, int slot) {
  return (double)*reinterpret_cast<const long long*>(stats + slot) / kPpFx;
}
__device__ __forceinline__ void pp_bbox_add(int* stats, int slot, bool on, int y, int x) {
  // one atomic quadruple per warp: shuffle-reduce the candidates first
  const unsigned m = __ballot_sync(0xffffffffu, on);
  if (!m) return;
  const int rmin = __reduce_min_sync(0xffffffffu, on ? y : INT_MAX), rmax = __reduce_max_sync(0xffffffffu, on ? y : -1);
  const int cmin = __reduce_min_sync(0xffffffffu, on ? x : INT_MAX), cmax = __reduce_max_sync(0xffffffffu, on ? x : -1);
  if ((threadIdx.x & 31) == 0) {
    atomicMin(stats + slot, rmin); atomicMax(stats + slot + 1, rmax);
    atomicMin(stats + slot + 2, cmin); atomicMax(stats + slot + 3, cmax);
  }
}
__device__ __forceinline__ void pp_count_add(int* stats, int slot, bool on) {
  const unsigned m = __ballot_sync(0xffffffffu, on);
  if (m && (threadIdx.x & 31) == 0) atomicAdd(stats + slot, __popc(m));
}

// tf.image.resize(x, [size, size]) source taps (bilinear, half-pixel centres, no antialias), float32 like the oracle:
// src = (i + 0.5) * (256 / size) - 0.5, clamped taps, no FMA contraction.
struct PpTap { int lo, hi; float f; };
__device__ __forceinline__ PpTap pp_tap(int i, float scale) {
  const float src = __fsub_rn(__fmul_rn((float)i + 0.5f, scale), 0.5f);
  const float lo = floorf(src);
  PpTap t;
  t.f = __fsub_rn(src, lo);
  const int l = (int)lo;
  t.lo = min(max(l, 0), PP_IMG - 1);
  t.hi = min(max(l + 1, 0), PP_IMG - 1);
  return t;
}
__device__ __forceinline__ float pp_lerp(float a, float b, float f) {
  return __fadd_rn(__fmul_rn(a, __fsub_rn(1.f, f)), __fmul_rn(b, f));
}
__device__ __forceinline__ float pp_mean3(float a, float b, float c) { return __fdiv_rn(__fadd_rn(__fadd_rn(a, b), c), 3.f); }

struct PpIn {
  const float *img, *gt, *rgb, *dif;     // [n][256][256][3|3|3|1]
  const unsigned char* masks;            // [n][7][256][256]
  const int* sizes;                      // [n]
};

// ---- K0: reset the statistics record and the component tables
__global__ void pp_reset_kernel(PpPlanes* planes) {
  const PpPlanes P = planes[blockIdx.y];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < PP_PIX) { P.csize[i] = 0; P.chair[i] = 0; }
  if (i < PS_WORDS) {
    int v = 0;
    if (i < 20) v = (i & 3) == 0 || (i & 3) == 2 ? INT_MAX : -1;
    P.stats[i] = v;
  }
}

// ---- K1: resize everything to [size, size], pad to 256 (:438-476); bounding boxes of the region masks (:480-488, 530, 562)
__global__ void pp_resize_kernel(PpIn in, PpPlanes* planes) {
  const int n = blockIdx.y;
  const PpPlanes P = planes[n];
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = p >> 8, x = p & 255;
  const int size = in.sizes[n];
  const bool inside = y < size && x < size;
  float v[10];
  unsigned char mk[PP_NMASK];
#pragma unroll
  for (int c = 0; c < 10; ++c) v[c] = 0.f;
#pragma unroll
  for (int k = 0; k < PP_NMASK; ++k) mk[k] = 0;
  if (inside) {
    const float scale = (float)(256.0 / (double)size);
    const PpTap ty = pp_tap(y, scale), tx = pp_tap(x, scale);
    const size_t b00 = (size_t)n * PP_PIX + ty.lo * PP_IMG + tx.lo, b01 = (size_t)n * PP_PIX + ty.lo * PP_IMG + tx.hi;
    const size_t b10 = (size_t)n * PP_PIX + ty.hi * PP_IMG + tx.lo, b11 = (size_t)n * PP_PIX + ty.hi * PP_IMG + tx.hi;
    auto samp = [&](const float* src, int C, int c) {
      const float a0 = pp_lerp(src[b00 * C + c], src[b10 * C + c], ty.f), a1 = pp_lerp(src[b01 * C + c], src[b11 * C + c], ty.f);
      return pp_lerp(a0, a1, tx.f);
    };
#pragma unroll
    for (int c = 0; c < 3; ++c) { v[c] = samp(in.img, 3, c); v[3 + c] = samp(in.gt, 3, c); v[6 + c] = samp(in.rgb, 3, c); }
    v[9] = samp(in.dif, 1, 0);
    const unsigned char* mb = in.masks + (size_t)n * PP_NMASK * PP_PIX;
#pragma unroll
    for (int k = 0; k < PP_NMASK; ++k) {
      const unsigned char* s = mb + (size_t)k * PP_PIX;
      const float a0 = pp_lerp((float)s[ty.lo * PP_IMG + tx.lo], (float)s[ty.hi * PP_IMG + tx.lo], ty.f);
      const float a1 = pp_lerp((float)s[ty.lo * PP_IMG + tx.hi], (float)s[ty.hi * PP_IMG + tx.hi], ty.f);
      mk[k] = (unsigned char)rintf(pp_lerp(a0, a1, tx.f));
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) { P.tmp[p * 3 + c] = v[c]; P.gt[p * 3 + c] = v[3 + c]; P.pred[p * 3 + c] = v[6 + c]; }
  P.mp[p] = v[9] * (float)mk[PP_FACE_HAIR];                                   // mask_pred * curr_mask (:476)
  P.inten[p] = pp_mean3(v[0], v[1], v[2]);
#pragma unroll
  for (int k = 0; k < PP_NMASK; ++k) P.masks[(size_t)k * PP_PIX + p] = mk[k];
  pp_bbox_add(P.stats, PS_NOSE, mk[PP_NOSE] == 1, y, x);
  pp_bbox_add(P.stats, PS_MOUTH, mk[PP_MOUTH] == 1, y, x);
  pp_bbox_add(P.stats, PS_BROW, mk[PP_EYEBROW] == 1, y, x);
  pp_bbox_add(P.stats, PS_FACE, mk[PP_FACE] == 1, y, x);
  pp_count_add(P.stats, PS_BROW_SUM, mk[PP_EYEBROW] == 1);
}

// ---- K2: mustache / mouth false positives (:479-496), forehead box (:528-535), mouth-and-below statistics (:541-549)
__global__ void pp_rules_kernel(PpPlanes* planes) {
  const PpPlanes P = planes[blockIdx.y];
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = p >> 8, x = p & 255;
  const int* S = P.stats;
  const bool have_nose = S[PS_NOSE + 1] >= 0, have_mouth = S[PS_MOUTH + 1] >= 0, have_brow = S[PS_BROW + 1] >= 0;
  float mp = P.mp[p];
  if (have_nose && have_mouth) {
    const int mid_nose = (int)((S[PS_NOSE + 1] + S[PS_NOSE]) / 2.0);
    const int um = S[PS_MOUTH], lm = S[PS_MOUTH + 1], left = S[PS_MOUTH + 2], right = S[PS_MOUTH + 3];
    if (y >= mid_nose && y < um && x >= left && x < right && mp < 0.018f) mp = 0.f;
    if (y >= um && y < lm && x >= left && x < right && mp < 0.02f) mp = 0.f;
    P.mp[p] = mp;
  }
  const bool face = P.masks[(size_t)PP_FACE * PP_PIX + p] == 1;
  // forehead: face pixels above the eyebrows
  const bool fh = have_brow && S[PS_BROW_SUM] * 3 > 30 && face && y < S[PS_BROW];
  pp_bbox_add(P.stats, PS_FOREHEAD, fh, y, x);
  if (have_mouth) {
    const bool roi = face && y >= S[PS_MOUTH];
    const bool sh = mp > 0.01f;
    pp_count_add(P.stats, PS_ROI_SUM, roi);
    pp_count_add(P.stats, PS_ROI_SHADOWED, roi && sh);
    // sum over pixels of mean_c(roi * tmp * shadowed): fixed point, one add per warp
    float mab = (roi && sh) ? pp_mean3(P.tmp[p * 3], P.tmp[p * 3 + 1], P.tmp[p * 3 + 2]) : 0.f;
    double w = (double)mab;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
    if ((threadIdx.x & 31) == 0 && w != 0.0) pp_fx_add(P.stats, PS_MAB_FX, w);
  }
}

// ---- K3: per-pixel threshold (:518-570), detected = mask_pred > threshold (:577), union-find initialisation
__global__ void pp_threshold_kernel(PpPlanes* planes) {
  const PpPlanes P = planes[blockIdx.y];
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = p >> 8, x = p & 255;
  const int* S = P.stats;
  const float mp = P.mp[p], inten = P.inten[p];
  const bool face = P.masks[(size_t)PP_FACE * PP_PIX + p] == 1, brow = P.masks[(size_t)PP_EYEBROW * PP_PIX + p] == 1;
  const int hair = (int)P.masks[(size_t)PP_FACE_HAIR * PP_PIX + p] - (int)P.masks[(size_t)PP_FACE * PP_PIX + p];
  float thr = 0.01f;
  if (hair > 0) thr = inten < 0.13f ? 0.004f : 0.02f;
  const bool have_brow = S[PS_BROW + 1] >= 0, have_mouth = S[PS_MOUTH + 1] >= 0, have_face = S[PS_FACE + 1] >= 0;
  if (have_brow && S[PS_BROW_SUM] * 3 > 30 && S[PS_FOREHEAD + 1] >= 0) {
    const int r0 = S[PS_FOREHEAD] + 20, r1 = S[PS_BROW] - 40, c0 = S[PS_FOREHEAD + 2] + 40, c1 = S[PS_FOREHEAD + 3] - 40;
    // numpy slice semantics of forehead_mask[r0:r1, c0:c1] (negative bounds wrap; they cannot occur for r0, c0 >= 20)
    const int rr1 = r1 < 0 ? r1 + PP_IMG : r1, cc1 = c1 < 0 ? c1 + PP_IMG : c1;
    if (y >= r0 && y < rr1 && x >= c0 && x < cc1 && inten < 0.4f) thr = -0.001f;
  }
  if (have_mouth) {
    const bool roi = face && y >= S[PS_MOUTH];
    const double frac = (double)((float)S[PS_ROI_SHADOWED] / (float)S[PS_ROI_SUM]);      // both counts x3 channels cancel
    const double mean_mab = pp_fx_get(S, PS_MAB_FX) / (double)S[PS_ROI_SHADOWED];
    bool kill = frac > 0.252 && frac < 0.268;
    kill = kill || (frac > 0.3 && frac < 0.31 && mean_mab > 0.358);
    kill = kill || (frac > 0.295 && frac < 0.3 && mean_mab > 0.22);
    if (kill && roi) thr = 1.0f;
  }
  if (have_brow && have_face && S[PS_BROW + 2] - S[PS_FACE + 2] == 0) {
    const double mid_face = S[PS_FACE + 2] * 0.8 + S[PS_FACE + 3] * 0.2;
    if (x < (int)mid_face && brow && inten > 0.1f) thr = 1.0f;
  }
  const bool det = mp > thr;
  P.detected[p] = det ? 1 : 0;
  P.label[p] = det ? p : -1;
}

// ---- K4: 4-connected components by union-find with atomicMin (label equivalence); cv2.connectedComponentsWithStats
__device__ __forceinline__ int pp_find(const int* label, int a) {
  while (true) {
    const int b = label[a];
    if (b == a) return a;
    a = b;
  }
}
__device__ __forceinline__ void pp_union(int* label, int a, int b) {
  while (true) {
    a = pp_find(label, a);
    b = pp_find(label, b);
    if (a == b) return;
    if (a > b) { const int t = a; a = b; b = t; }
    const int old = atomicMin(label + b, a);      // hook the larger root under the smaller one
    if (old == b) return;
    b = old;
  }
}
__global__ void pp_ccl_merge_kernel(PpPlanes* planes) {
  const PpPlanes P = planes[blockIdx.y];
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (!P.detected[p]) return;
  const int y = p >> 8, x = p & 255;
  if (x + 1 < PP_IMG && P.detected[p + 1]) pp_union(P.label, p, p + 1);
  if (y + 1 < PP_IMG && P.detected[p + PP_IMG]) pp_union(P.label, p, p + PP_IMG);
}
__global__ void pp_ccl_flatten_kernel(PpPlanes* planes) {
  const PpPlanes P = planes[blockIdx.y];
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (!P.detected[p]) return;
  const int r = pp_find(P.label, p);
  P.label[p] = r;         // roots keep label[r] == r, so concurrent finds stay correct
  atomicAdd(P.csize + r, 1);
  const int hair = (int)P.masks[(size_t)PP_FACE_HAIR * PP_PIX + p] - (int)P.masks[(size_t)PP_FACE * PP_PIX + p];
  if (hair) atomicAdd(P.chair + r, hair);
}
__global__ void pp_cc_max_kernel(PpPlanes* planes) {
  const PpPlanes P = planes[blockIdx.y];
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int sz = (P.detected[p] && P.label[p] == p) ? P.csize[p] : 0;
  const int m = __reduce_max_sync(0xffffffffu, sz);
  if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(P.stats + PS_CC_MAX, m);
}

// ---- K5: keep components with size >= 0.45 max and hair fraction < 0.8 (:598-611); nose statistics (:650-652)
__global__ void pp_select_kernel(PpPlanes* planes) {
  const PpPlanes P = planes[blockIdx.y];
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  bool keep = false;
  if (P.detected[p]) {
    const int r = P.label[p];
    const double sz = (double)P.csize[r];
    keep = sz >= 0.45 * (double)P.stats[PS_CC_MAX] && (double)P.chair[r] / sz < 0.8;
  }
  P.img2[p] = keep ? 1 : 0;
  const bool nose = P.masks[(size_t)PP_NOSE * PP_PIX + p] == 1;
  const float inten = P.inten[p];
  pp_count_add(P.stats, PS_IMG2_SUM, keep);
  pp_count_add(P.stats, PS_NOSE_SUM, nose);
  pp_count_add(P.stats, PS_NOSE_SHADOW, nose && keep && inten > 0.f);
  double w = keep ? (double)inten : 0.0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
  if ((threadIdx.x & 31) == 0 && w != 0.0) pp_fx_add(P.stats, PS_SHADOW_FX, w);
}

// ---- K6: nose rule (:653-662), composite + clip (:711, 718), squared error for the PSNR (:725)
__global__ void pp_final_kernel(PpPlanes* planes, float* __restrict__ final_out, float* __restrict__ detected_out) {
  const int n = blockIdx.y;
  const PpPlanes P = planes[n];
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = p >> 8, x = p & 255;
  const int* S = P.stats;
  bool det = P.img2[p] != 0;
  if (S[PS_NOSE + 1] >= 0 && S[PS_IMG2_SUM] > 0) {
    const double mean_intensity = pp_fx_get(S, PS_SHADOW_FX) / (double)S[PS_IMG2_SUM];
    const double frac = (double)S[PS_NOSE_SHADOW] / (double)S[PS_NOSE_SUM];
    if ((frac > 0.15 && frac < 0.25) || (frac > 0.30 && frac < 0.31) || (frac > 0.34 && frac < 0.35)) {
      const double mid_h = (S[PS_NOSE + 1] + S[PS_NOSE]) / 2.0, mid_w = (S[PS_NOSE + 3] + S[PS_NOSE + 2]) / 2.0;
      const int r0 = (int)mid_h, r1 = S[PS_NOSE + 1] + (mean_intensity < 0.15 ? 5 : 65);
      const int c0 = (int)(mid_w - 35.0), c1 = (int)(mid_w + 35.0);
      const int cc0 = c0 < 0 ? max(c0 + PP_IMG, 0) : c0;               // numpy slice semantics for a negative start
      if (y >= r0 && y < r1 && x >= cc0 && x < c1) det = false;
    }
  }
  const float d = det ? 1.f : 0.f;
  if (detected_out) detected_out[(size_t)n * PP_PIX + p] = d;
  double se = 0.0;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float v = fminf(fmaxf(__fadd_rn(__fmul_rn(P.pred[p * 3 + c], d), __fmul_rn(P.tmp[p * 3 + c], 1.f - d)), 0.f), 1.f);
    final_out[((size_t)n * PP_PIX + p) * 3 + c] = v;
    P.pred[p * 3 + c] = v;                      // the SSIM kernel reads the final image from here
    const double e = (double)P.gt[p * 3 + c] - (double)v;
    se += e * e;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
  if ((threadIdx.x & 31) == 0 && se != 0.0) pp_fx_add(P.stats, PS_SQERR_FX, se);
}

// ---- K7: tf.image.ssim(gt, final, max_val = 1): 11 x 11 Gaussian (sigma 1.5), VALID, k1 = .01, k2 = .03, mean of
// luminance * contrast-structure over the 246 x 246 positions and the 3 channels (:724).  One block = 16 x 16 outputs
// of one channel; separable filter in shared memory, float64.
__global__ void __launch_bounds__(256) pp_ssim_kernel(PpPlanes* planes) {
  const PpPlanes P = planes[blockIdx.z];
  __shared__ float sa[26][27], sb[26][27];
  __shared__ double hq[5][26][16];
  __shared__ double g[11];
  const int c = blockIdx.y, tiles = 16;                       // 16 x 16 tiles of 16 x 16 cover 246 x 246
  const int ty0 = (blockIdx.x / tiles) * 16, tx0 = (blockIdx.x % tiles) * 16;
  const int t = threadIdx.x;
  if (t < 11) {
    double s = 0.0;
    for (int i = 0; i < 11; ++i) s += exp(-((i - 5.0) * (i - 5.0)) / (2.0 * 1.5 * 1.5));
    g[t] = exp(-((t - 5.0) * (t - 5.0)) / (2.0 * 1.5 * 1.5)) / s;
  }
  for (int i = t; i < 26 * 26; i += 256) {
    const int yy = ty0 + i / 26, xx = tx0 + i % 26;
    const bool ok = yy < PP_IMG && xx < PP_IMG;
    sa[i / 26][i % 26] = ok ? P.gt[(yy * PP_IMG + xx) * 3 + c] : 0.f;
    sb[i / 26][i % 26] = ok ? P.pred[(yy * PP_IMG + xx) * 3 + c] : 0.f;
  }
  __syncthreads();
  for (int i = t; i < 26 * 16; i += 256) {                    // horizontal pass
    const int r = i / 16, x = i % 16;
    double q[5] = {0, 0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const double a = sa[r][x + k], b = sb[r][x + k], w = g[k];
      q[0] += w * a; q[1] += w * b; q[2] += w * a * a; q[3] += w * b * b; q[4] += w * a * b;
    }
#pragma unroll
    for (int j = 0; j < 5; ++j) hq[j][r][x] = q[j];
  }
  __syncthreads();
  const int oy = t / 16, ox = t % 16;
  double val = 0.0;
  if (ty0 + oy < PP_IMG - 10 && tx0 + ox < PP_IMG - 10) {
    double q[5] = {0, 0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < 11; ++k)
#pragma unroll
      for (int j = 0; j < 5; ++j) q[j] += g[k] * hq[j][oy + k][ox];
    const double c1 = 0.01 * 0.01, c2 = 0.03 * 0.03;
    const double va = q[2] - q[0] * q[0], vb = q[3] - q[1] * q[1], cov = q[4] - q[0] * q[1];
    val = ((2.0 * q[0] * q[1] + c1) / (q[0] * q[0] + q[1] * q[1] + c1)) * ((2.0 * cov + c2) / (va + vb + c2));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
  if ((t & 31) == 0 && val != 0.0) pp_fx_add(P.stats, PS_SSIM_FX, val);
}

// ---- K8: metrics[n] = (ssim, psnr)
__global__ void pp_metrics_kernel(PpPlanes* planes, float* __restrict__ metrics, int n_img) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_img) return;
  const int* S = planes[n].stats;
  const double ssim = pp_fx_get(S, PS_SSIM_FX) / (246.0 * 246.0 * 3.0);
  const double mse = pp_fx_get(S, PS_SQERR_FX) / (double)(PP_PIX * 3);
  metrics[2 * n] = (float)ssim;
  metrics[2 * n + 1] = (float)(mse > 0.0 ? 10.0 * log10(1.0 / mse) : INFINITY);
}

constexpr size_t kPpBytesPerImage = (size_t)PP_PIX * (3 * 3 * 4 + 2 * 4 + PP_NMASK + 2 + 3 * 4) + PS_WORDS * 4 + 256;

}  // namespace bsr
