// Bandwidth-bound glue between the convolutions: 256->32 resizes, grey composition, hole mask,
// temporal sharing (ShareLayer), colour tail, caller-side blends.  T = activation storage type
// (h16 in the product path, float in FP32CHECK).  Reference lines are cited per kernel.
#pragma once
#include "common.cuh"

namespace bsr {

constexpr int IMG = 256, FEAT = 32, CELL = 8;

// tf.image.resize(x, [32,32]) of a 256x256 map (bilinear, half-pixel centres, no antialias): source
// coordinate 8i+3.5 -> rows/cols (8i+3, 8i+4) with weights 1/2 (model.py:237, 256; warp.py:137).
__device__ __forceinline__ float resize8(const float* __restrict__ p, int n, int i, int j, int ld, int c) {
  const float* b = p + (((size_t)n * IMG + (CELL * i + 3)) * IMG + (CELL * j + 3)) * ld + c;
  float top = b[0] * 0.5f + b[ld] * 0.5f;
  float bot = b[(size_t)IMG * ld] * 0.5f + b[(size_t)IMG * ld + ld] * 0.5f;
  return top * 0.5f + bot * 0.5f;
}

// uv[N,256,256,3] -> uvs[N,32,32,3]   (model.py:237)
__global__ void uv_small_kernel(const float* __restrict__ uv, float* __restrict__ uvs, int n_img) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_img * FEAT * FEAT * 3) return;
  int c = idx % 3, cell = idx / 3;
  int j = cell % FEAT, i = (cell / FEAT) % FEAT, n = cell / (FEAT * FEAT);
  uvs[idx] = resize8(uv, n, i, j, 3, c);
}

// conv1 input packing: img fp32 [N,256,256,3] -> 16-bit [N][257][264][8].  Packed row yp holds image rows yp-1 and yp,
// packed pixel xp holds image pixel xp-3: 8 channels = [R G B 0 of row yp-1 | R G B 0 of row yp] (zero outside the
// image).  The 7-tap window of output x in packed row yp starts at xp = x and is ONE 64-element K block that covers
// TWO filter rows, so a 7x7 tile needs 7 row-pair steps instead of 14 row steps (K utilisation 42/64 instead of 21/64).
__global__ void pack_img_kernel(const float* __restrict__ img, h16* __restrict__ out, long long n_img) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_img * (IMG + 1) * (IMG + 8)) return;
  const int xp = (int)(idx % (IMG + 8));
  const long long t = idx / (IMG + 8);
  const int yp = (int)(t % (IMG + 1));
  const long long n = t / (IMG + 1);
  const int x = xp - 3;
  uint4 o = make_uint4(0u, 0u, 0u, 0u);
  if (x >= 0 && x < IMG) {
    if (yp >= 1) {
      const float* s = img + ((n * IMG + (yp - 1)) * IMG + x) * 3;
      o.x = pack_h16x2(s[0], s[1]);
      o.y = pack_h16x2(s[2], 0.f);
    }
    if (yp < IMG) {
      const float* s = img + ((n * IMG + yp) * IMG + x) * 3;
      o.z = pack_h16x2(s[0], s[1]);
      o.w = pack_h16x2(s[2], 0.f);
    }
  }
  reinterpret_cast<uint4*>(out)[idx] = o;
}

// reg[N,256,256,6] -> off[N,32,32,4] = 32 * resize(reg)[(0,1) of reg_in, (0,1) of reg_out]
// (model_with_TSM.py:207; warp.py:137-139).  NaN offsets (generate_offset_map does no nan_to_num,
// warp.py:194-213) are defined as 0 here.
__global__ void reg_small_kernel(const float* __restrict__ reg, float* __restrict__ off, int n_img) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_img * FEAT * FEAT * 4) return;
  int k = idx & 3, cell = idx >> 2;
  int j = cell % FEAT, i = (cell / FEAT) % FEAT, n = cell / (FEAT * FEAT);
  int c = (k < 2) ? k : k + 1;          // channels 0,1,3,4
  float v = resize8(reg, n, i, j, 6, c) * (float)FEAT;
  off[idx] = (v == v) ? v : 0.f;
}

// Host path: only image rows 8i+3 and 8i+4 of uv / reg ever reach the model (resize8 above), so the pipelined host path
// uploads just those rows (one strided 2-D DMA, a quarter of the bytes): rows[n][32][2][256][C].  Same taps, same
// operation order as resize8 -> bit-identical results.
__device__ __forceinline__ float resize8_rows(const float* __restrict__ rows, int n, int i, int j, int C, int c) {
  const float* b = rows + ((((size_t)n * FEAT + i) * 2) * IMG + (CELL * j + 3)) * C + c;
  float top = b[0] * 0.5f + b[C] * 0.5f;
  float bot = b[(size_t)IMG * C] * 0.5f + b[(size_t)IMG * C + C] * 0.5f;
  return top * 0.5f + bot * 0.5f;
}
__global__ void uv_rows_small_kernel(const float* __restrict__ rows, float* __restrict__ uvs, int n_img) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_img * FEAT * FEAT * 3) return;
  int c = idx % 3, cell = idx / 3;
  int j = cell % FEAT, i = (cell / FEAT) % FEAT, n = cell / (FEAT * FEAT);
  uvs[idx] = resize8_rows(rows, n, i, j, 3, c);
}
__global__ void reg_rows_small_kernel(const float* __restrict__ rows, float* __restrict__ off, int n_img) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_img * FEAT * FEAT * 4) return;
  int k = idx & 3, cell = idx >> 2;
  int j = cell % FEAT, i = (cell / FEAT) % FEAT, n = cell / (FEAT * FEAT);
  int c = (k < 2) ? k : k + 1;          // channels 0,1,3,4
  float v = resize8_rows(rows, n, i, j, 6, c) * (float)FEAT;
  off[idx] = (v == v) ? v : 0.f;
}

// ---- compact host I/O -------------------------------------------------------
// img_u8 -> fp32 `/ 255.` (dataset.py:119,159).  16 bytes in, 64 bytes out per thread; n16 = bytes/16.
__global__ void expand_u8_kernel(const uint4* __restrict__ in, float4* __restrict__ out, long long n16) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n16) return;
  uint4 v = __ldg(in + idx);
  uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float4 o;
    o.x = __fdiv_rn((float)(w[k] & 0xffu), 255.f);
    o.y = __fdiv_rn((float)((w[k] >> 8) & 0xffu), 255.f);
    o.z = __fdiv_rn((float)((w[k] >> 16) & 0xffu), 255.f);
    o.w = __fdiv_rn((float)(w[k] >> 24), 255.f);
    out[idx * 4 + k] = o;
  }
}

// reg32[N,32,32,6] (already resized) -> off[N,32,32,4], same selection/scale/NaN rule as reg_small_kernel.
__global__ void reg32_off_kernel(const float* __restrict__ reg32, float* __restrict__ off, int n_img) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_img * FEAT * FEAT * 4) return;
  int k = idx & 3, cell = idx >> 2;
  int c = (k < 2) ? k : k + 1;
  float v = reg32[(size_t)cell * 6 + c] * (float)FEAT;
  off[idx] = (v == v) ? v : 0.f;
}

// rgb fp32 -> uint8 rint(clip(x,0,1)*255) (train_test_GSC.py:809; utils.py:221,180): 16 floats -> 16 bytes.
__global__ void rgb_to_u8_kernel(const float4* __restrict__ in, uint4* __restrict__ out, long long n16) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n16) return;
  uint32_t w[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float4 v = __ldg(in + idx * 4 + k);
    auto q = [](float x) { return (uint32_t)__float2int_rn(fminf(fmaxf(x, 0.f), 1.f) * 255.f); };
    w[k] = q(v.x) | (q(v.y) << 8) | (q(v.z) << 16) | (q(v.w) << 24);
  }
  out[idx] = make_uint4(w[0], w[1], w[2], w[3]);
}

// dif fp32 -> binary16 (round to nearest even): 8 floats -> 16 bytes.
__global__ void f32_to_f16_kernel(const float4* __restrict__ in, uint4* __restrict__ out, long long n8) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n8) return;
  float4 a = __ldg(in + idx * 2), b = __ldg(in + idx * 2 + 1);
  __half2 h0 = __floats2half2_rn(a.x, a.y), h1 = __floats2half2_rn(a.z, a.w);
  __half2 h2 = __floats2half2_rn(b.x, b.y), h3 = __floats2half2_rn(b.z, b.w);
  out[idx] = make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1),
                        *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
}

// First-half res-stack input: write uv at channel uv_off and zero [z0, z1)   (model.py:238;
// model_with_TSM.py:272).  One thread per (pixel, channel of the written range).
template <typename T>
__global__ void assemble_uv_kernel(T* __restrict__ x, int ld, const float* __restrict__ uvs, int uv_off,
                                   int z0, int z1, int n_pix) {
  int span = 3 + (z1 - z0);
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)n_pix * span) return;
  int k = (int)(idx % span);
  size_t pix = (size_t)(idx / span);
  if (k < 3) stf<T>(x, pix * ld + uv_off + k, uvs[pix * 3 + k]);
  else stf<T>(x, pix * ld + z0 + (k - 3), 0.f);
}

// h16 product-path version of assemble_uv_kernel: one thread per 16-byte (8-channel) chunk of the written range
// [uv_off & ~7, ld): channels [uv_off, uv_off+3) = uv, everything above = 0, channels below uv_off are preserved.
__global__ void assemble_uv_vec_kernel(h16* __restrict__ x, int ld, const float* __restrict__ uvs, int uv_off, int n_pix) {
  const int k0 = uv_off >> 3, nk = (ld >> 3) - k0;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)n_pix * nk) return;
  const int k = k0 + (int)(idx % nk);
  const size_t pix = (size_t)(idx / nk);
  uint4* dst = reinterpret_cast<uint4*>(x + pix * ld + 8 * k);
  uint4 w = make_uint4(0u, 0u, 0u, 0u);
  if (8 * k < uv_off + 3) {                      // chunk holds preserved and / or uv channels
    if (8 * k < uv_off) w = *dst;
    h16* e = reinterpret_cast<h16*>(&w);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = 8 * k + i;
      if (c >= uv_off + 3) e[i] = f32_to_h16(0.f);
      else if (c >= uv_off) e[i] = f32_to_h16(uvs[pix * 3 + (c - uv_off)]);
    }
  }
  *dst = w;
}

// model.py:246-252: mask = tanh(conv2(y)); con = conv3(y); gs = grey*(1+mask)+con; dif = gs-grey;
// mask22 = [relu(mask), 0, relu(-mask)].  raw[...,0] = conv2(y), raw[...,1] = conv3(y) (bias included).
// Also writes gs as channel `gs_c` of the clr_conv1 input buffer and zeroes its pad channels.
template <typename T>
__global__ void compose_kernel(const float* __restrict__ raw, const float* __restrict__ img,
                               float* __restrict__ gs_out, float* __restrict__ mask22_out,
                               float* __restrict__ difgs, T* __restrict__ cat1, int cat_ld, int gs_c,
                               long long n_pix) {
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pix) return;
  float2 r = *reinterpret_cast<const float2*>(raw + 2 * p);
  float mask = tanhf(r.x), con = r.y;
  float g = img[3 * p] * kGrayR + img[3 * p + 1] * kGrayG + img[3 * p + 2] * kGrayB;
  float gs = g * (1.f + mask) + con;
  difgs[p] = gs - g;
  if (gs_out) gs_out[p] = gs;
  if (mask22_out) {
    mask22_out[3 * p] = fmaxf(mask, 0.f);
    mask22_out[3 * p + 1] = mask * 0.f;
    mask22_out[3 * p + 2] = fmaxf(-mask, 0.f);
  }
  stf<T>(cat1, (size_t)p * cat_ld + gs_c, gs);
  for (int c = gs_c + 1; c < cat_ld; ++c) stf<T>(cat1, (size_t)p * cat_ld + c, 0.f);
}

// model.py:256-259 / model_with_TSM.py:290-294: bmask = resize(dif,[32,32]) > 0.1 (strict);
// x_hole = x*(1-bmask); out = [x_hole (cx ch) | bmask | ... | uv at uv_off]; zero [z0,z1).
// One warp per 32x32 cell.
template <typename T>
__global__ void hole_kernel(const float* __restrict__ difgs, const T* xa, int lda, T* xb, int ldb, int cx, const float* __restrict__ uvs, int uv_off,
                            int z0, int z1, float* __restrict__ bmask_out, float* __restrict__ difsmall_out,
                            int n_cells) {
  int cell = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (cell >= n_cells) return;
  int j = cell % FEAT, i = (cell / FEAT) % FEAT, n = cell / (FEAT * FEAT);
  float d = resize8(difgs, n, i, j, 1, 0);
  float bm = d > kHoleThr ? 1.f : 0.f;
  float keep = 1.f - bm;
  const T* src = xa + (size_t)cell * lda;
  T* dst = xb + (size_t)cell * ldb;
  int c_done = 0;
  if (src == dst) {
    // in place (both halves of the network use the same channel stride: GSC): a kept cell is already what it has to be,
    // a masked cell is cleared without being read
    if (bm != 0.f) {
      if (sizeof(T) == 2 && (ldb & 7) == 0) {
        c_done = cx & ~7;
        for (int c = lane * 8; c < c_done; c += 256) *reinterpret_cast<uint4*>(dst + c) = make_uint4(0u, 0u, 0u, 0u);
      }
      for (int c = c_done + lane; c < cx; c += 32) stf<T>(dst, c, 0.f);
    }
  } else {
    if (sizeof(T) == 2 && (lda & 7) == 0 && (ldb & 7) == 0) {
      // 16-bit storage: 8 channels (16 bytes) per lane and step; keep is 0 or 1, so the product is a select
      c_done = cx & ~7;
      for (int c = lane * 8; c < c_done; c += 256) {
        uint4 v = *reinterpret_cast<const uint4*>(src + c);
        if (bm != 0.f) v = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(dst + c) = v;
      }
    }
    for (int c = c_done + lane; c < cx; c += 32) stf<T>(dst, c, ldf<T>(src, c) * keep);
  }
  if (lane == 0) {
    stf<T>(dst, cx, bm);
    bmask_out[cell] = bm;
    difsmall_out[cell] = d;
  }
  if (lane < 3) stf<T>(dst, uv_off + lane, uvs[(size_t)cell * 3 + lane]);
  for (int c = z0 + lane; c < z1; c += 32) stf<T>(dst, c, 0.f);
}

// In-place form of hole_kernel for 16-bit storage (GSC: both halves of the network share the channel stride): one THREAD
// per cell takes the four resize taps and writes the cell's few side values; the warp then clears the cx feature channels
// of its masked cells together (16 bytes per lane), kept cells are not touched.  Needs ld % 8 == 0.
__global__ void __launch_bounds__(256) hole_inplace_h16_kernel(const float* __restrict__ difgs, h16* __restrict__ xb, int ld, int cx,
                                                               const float* __restrict__ uvs, int uv_off, int z0, int z1,
                                                               float* __restrict__ bmask_out, float* __restrict__ difsmall_out,
                                                               int n_cells) {
  const int cell = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
  const bool live = cell < n_cells;
  bool masked = false;
  if (live) {
    const int j = cell % FEAT, i = (cell / FEAT) % FEAT, n = cell / (FEAT * FEAT);
    const float d = resize8(difgs, n, i, j, 1, 0);
    masked = d > kHoleThr;
    h16* dst = xb + (size_t)cell * ld;
    dst[cx] = f32_to_h16(masked ? 1.f : 0.f);
    bmask_out[cell] = masked ? 1.f : 0.f;
    difsmall_out[cell] = d;
#pragma unroll
    for (int k = 0; k < 3; ++k) dst[uv_off + k] = f32_to_h16(uvs[(size_t)cell * 3 + k]);
    for (int c = z0; c < z1; ++c) dst[c] = f32_to_h16(0.f);
  }
  unsigned m = __ballot_sync(0xffffffffu, masked);
  const int cell0 = cell - lane, c_vec = cx & ~7;
  while (m) {
    const int src_lane = __ffs(m) - 1;
    m &= m - 1;
    h16* dst = xb + (size_t)(cell0 + src_lane) * ld;
    for (int c = lane * 8; c < c_vec; c += 256) *reinterpret_cast<uint4*>(dst + c) = make_uint4(0u, 0u, 0u, 0u);
    if (lane < cx - c_vec) dst[c_vec + lane] = f32_to_h16(0.f);
  }
}

// ---- ShareLayer (model_with_TSM.py:204-229) ---------------------------------------------------
// tf_batch_map_coordinates (warp.py:71-115): clip coords to [0, s-1], corners floor/ceil,
// v_t = lt + (rt-lt)*o0 ; v_b = lb + (rb-lb)*o0 ; out = v_t + (v_b-v_t)*o1, with coordinate 0 = row.
struct WarpTap { int lt, rt, lb, rb; float o0, o1; };
__device__ __forceinline__ WarpTap warp_tap(float off0, float off1, int i, int j) {
  float c0 = fminf(fmaxf(off0 + (float)i, 0.f), (float)(FEAT - 1));
  float c1 = fminf(fmaxf(off1 + (float)j, 0.f), (float)(FEAT - 1));
  float f0 = floorf(c0), f1 = floorf(c1);
  int r0 = (int)f0, r1 = (int)ceilf(c0), q0 = (int)f1, q1 = (int)ceilf(c1);
  WarpTap w;
  w.lt = r0 * FEAT + q0;      // (floor0, floor1)
  w.rb = r1 * FEAT + q1;      // (ceil0, ceil1)
  w.lb = r0 * FEAT + q1;      // warp.py:88 (lt[0], rb[1])
  w.rt = r1 * FEAT + q0;      // warp.py:89 (rb[0], lt[1])
  w.o0 = c0 - f0;
  w.o1 = c1 - f1;
  return w;
}
__device__ __forceinline__ float warp_mix(float lt, float rt, float lb, float rb, float o0, float o1) {
  float vt = lt + (rt - lt) * o0;
  float vb = lb + (rb - lb) * o0;
  return vt + (vb - vt) * o1;
}

// 4 consecutive channels per thread: 8-byte (16-bit storage) / 16-byte (fp32) accesses, fully coalesced across a warp.
template <typename T> __device__ __forceinline__ void ld4(const T* p, float* v);
template <> __device__ __forceinline__ void ld4<float>(const float* p, float* v) {
  const float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <> __device__ __forceinline__ void ld4<h16>(const h16* p, float* v) {
  const uint2 t = *reinterpret_cast<const uint2*>(p);
  const float2 a = unpack_h16x2(t.x), b = unpack_h16x2(t.y);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
template <typename T> __device__ __forceinline__ void st4(T* p, const float* v);
template <> __device__ __forceinline__ void st4<float>(float* p, const float* v) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <> __device__ __forceinline__ void st4<h16>(h16* p, const float* v) {
  *reinterpret_cast<uint2*>(p) = make_uint2(pack_h16x2(v[0], v[1]), pack_h16x2(v[2], v[3]));
}

// ShareLayer, first half (model_with_TSM.py:204-222): sh[chunk][pix][0:C] = max_f warp_in(x_f), sh[chunk][pix][C:2C] =
// mean_f warp_in(x_f), stored in the activation type with row stride ldsh (a multiple of 4).  One WARP per (chunk,
// pixel): the bilinear taps are warp-uniform, the lanes stride over 4-channel vectors (8-byte / 16-byte accesses, 256 /
// 512 contiguous bytes per warp instruction); the frames of a chunk are reduced in registers (F = 2 or 10 - the
// reduction axis is a loop of every lane, so no cross-lane step is needed; hoisting the tap computation out of the
// vector loop with per-iteration accumulators measured SLOWER: 145 vs 124 us for 128 frames of 291 channels).  Needs ld % 4 == 0.
template <typename T>
__global__ void share_reduce_kernel(const T* __restrict__ x, int ld, int C, const float* __restrict__ off, int frame,
                                    T* __restrict__ sh, int ldsh, int n_cells) {
  const int cell = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (cell >= n_cells) return;
  const int pix = cell % (FEAT * FEAT);
  const size_t chunk = (size_t)(cell / (FEAT * FEAT));
  const int i = pix / FEAT, j = pix % FEAT, Cv = (C + 3) >> 2;
  T* d = sh + (chunk * FEAT * FEAT + pix) * ldsh;
  const float fr = (float)frame;
  for (int cv = lane; cv < Cv; cv += 32) {
    const int c0 = 4 * cv;
    float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY}, sum[4] = {0.f, 0.f, 0.f, 0.f};
    for (int f = 0; f < frame; ++f) {
      const size_t n = chunk * frame + f;
      const float2 o = *reinterpret_cast<const float2*>(off + (n * FEAT * FEAT + pix) * 4);       // warp-uniform
      const WarpTap w = warp_tap(o.x, o.y, i, j);
      const T* b = x + n * FEAT * FEAT * ld + c0;
      float lt[4], rt[4], lb[4], rb[4];
      ld4<T>(b + (size_t)w.lt * ld, lt);
      ld4<T>(b + (size_t)w.rt * ld, rt);
      ld4<T>(b + (size_t)w.lb * ld, lb);
      ld4<T>(b + (size_t)w.rb * ld, rb);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float v = warp_mix(lt[k], rt[k], lb[k], rb[k], w.o0, w.o1);
        mx[k] = fmaxf(mx[k], v);
        sum[k] += v;
      }
    }
    if (c0 + 4 <= C) {
      st4<T>(d + c0, mx);
      if ((C & 3) == 0) {
        const float mean[4] = {sum[0] / fr, sum[1] / fr, sum[2] / fr, sum[3] / fr};
        st4<T>(d + C + c0, mean);
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) stf<T>(d, C + c0 + k, sum[k] / fr);
      }
    } else {
      for (int k = 0; c0 + k < C; ++k) {
        stf<T>(d, c0 + k, mx[k]);
        stf<T>(d, C + c0 + k, sum[k] / fr);
      }
    }
  }
}

// ShareLayer, second half (model_with_TSM.py:223-226): x[n][pix][coff + c] = warp_out(sh[chunk])[pix][c], c < 2C.
// One warp per (frame, pixel), lanes over 4-channel vectors; coff % 4 == 0, ld % 4 == 0, ldsh % 4 == 0.
template <typename T>
__global__ void share_out_kernel(const T* __restrict__ sh, int ldsh, int C2, const float* __restrict__ off, int frame,
                                 T* __restrict__ x, int ld, int coff, int n_cells) {
  const int cell = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (cell >= n_cells) return;
  const int pix = cell % (FEAT * FEAT);
  const size_t n = (size_t)(cell / (FEAT * FEAT));
  const size_t chunk = n / frame;
  const int i = pix / FEAT, j = pix % FEAT, Qv = (C2 + 3) >> 2;
  const float2 o = *reinterpret_cast<const float2*>(off + (n * FEAT * FEAT + pix) * 4 + 2);
  const WarpTap w = warp_tap(o.x, o.y, i, j);
  const T* b = sh + chunk * FEAT * FEAT * ldsh;
  T* d = x + (n * FEAT * FEAT + pix) * ld + coff;
  for (int qv = lane; qv < Qv; qv += 32) {
    const int c0 = 4 * qv;
    float lt[4], rt[4], lb[4], rb[4], v[4];
    ld4<T>(b + (size_t)w.lt * ldsh + c0, lt);
    ld4<T>(b + (size_t)w.rt * ldsh + c0, rt);
    ld4<T>(b + (size_t)w.lb * ldsh + c0, lb);
    ld4<T>(b + (size_t)w.rb * ldsh + c0, rb);
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = warp_mix(lt[k], rt[k], lb[k], rb[k], w.o0, w.o1);
    if (c0 + 4 <= C2) {
      st4<T>(d + c0, v);
    } else {
      for (int k = 0; c0 + k < C2; ++k) stf<T>(d, c0 + k, v[k]);
    }
  }
}

// ---- 16-bit ShareLayer kernels (the product path).  The warp-per-cell kernels above are bound by instruction issue, not
// by memory (ncu: 2.6-3.1 IPC, 13 % of DRAM peak): tap arithmetic repeated per vector and frame, scalar fp32 blends, 8-byte
// accesses.  Here
//   * the bilinear taps of every (frame, pixel, direction) are computed ONCE per forward into 16-byte records;
//   * eight lanes share a cell (a warp works on four cells), each lane moves 16-byte vectors of 8 channels;
//   * the blend runs on packed fp32 FMAs (fma.rn.f32x2; the same operations in the same order as warp_mix);
//   * the shared features are stored in the alignment of their DESTINATION: row = [shift unused | max: C | mean: C] with
//     shift = coff % 8, so the un-warp kernel reads and writes aligned 16-byte vectors only and the reduce kernel (which
//     has 2 x frame fewer vectors to store) takes the odd alignments (C = 291 is odd) on its store side.
// Needs ld % 8 == 0 and coff % 4 == 0.
struct __align__(16) TapRec { uint32_t lt_rt, lb_rb; float o0, o1; };     // pixel indices (< 1024) in 16-bit halves

// off[n][pix][4] = (in-warp offset, out-warp offset) -> taps[n][pix][2]
__global__ void share_taps_kernel(const float* __restrict__ off, TapRec* __restrict__ taps, int n_rec) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_rec) return;
  const int pix = (idx >> 1) & (FEAT * FEAT - 1);
  const float2 o = *reinterpret_cast<const float2*>(off + (size_t)idx * 2);
  const WarpTap w = warp_tap(o.x, o.y, pix / FEAT, pix % FEAT);
  TapRec r;
  r.lt_rt = (uint32_t)w.lt | ((uint32_t)w.rt << 16);
  r.lb_rb = (uint32_t)w.lb | ((uint32_t)w.rb << 16);
  r.o0 = w.o0;
  r.o1 = w.o1;
  taps[idx] = r;
}

__device__ __forceinline__ uint64_t f2_pack(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t h2_to_f2(uint32_t w) {
  const float2 f = unpack_h16x2(w);
  return f2_pack(f.x, f.y);
}
// warp_mix on two channels at once: fma(x, -1, y) is the correctly rounded y - x, the other steps are the FMAs the scalar
// form contracts to
__device__ __forceinline__ uint64_t warp_mix2(uint32_t lt, uint32_t rt, uint32_t lb, uint32_t rb, uint64_t o0, uint64_t o1,
                                              uint64_t m1) {
  const uint64_t a = h2_to_f2(lt), b = h2_to_f2(rt), c = h2_to_f2(lb), d = h2_to_f2(rb);
  const uint64_t vt = f2_fma(f2_fma(a, m1, b), o0, a);
  const uint64_t vb = f2_fma(f2_fma(c, m1, d), o0, c);
  return f2_fma(f2_fma(vt, m1, vb), o1, vt);
}
// 8 consecutive 16-bit values (4 words) to a destination of ANY 2-byte alignment; the alignment is the same for every
// thread of a launch, so the branches do not diverge
__device__ __forceinline__ void store8_any(h16* dst, const uint32_t* w) {
  const unsigned a = ((unsigned)(uintptr_t)dst >> 1) & 7u;
  if (a == 0) {
    *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
  } else if ((a & 3u) == 0) {
    *reinterpret_cast<uint2*>(dst) = make_uint2(w[0], w[1]);
    *reinterpret_cast<uint2*>(dst + 4) = make_uint2(w[2], w[3]);
  } else if ((a & 1u) == 0) {
#pragma unroll
    for (int k = 0; k < 4; ++k) *reinterpret_cast<uint32_t*>(dst + 2 * k) = w[k];
  } else {
    unsigned short* d16 = reinterpret_cast<unsigned short*>(dst);
    d16[0] = (unsigned short)(w[0] & 0xffffu);
#pragma unroll
    for (int k = 0; k < 3; ++k) *reinterpret_cast<uint32_t*>(dst + 1 + 2 * k) = __funnelshift_r(w[k], w[k + 1], 16);
    d16[7] = (unsigned short)(w[3] >> 16);
  }
}

// First half (model_with_TSM.py:204-222): sh[chunk][pix][shift + c] = max_f warp_in(x_f)[c], [shift + C + c] = mean_f.
// The eight lanes of a cell assemble the row in shared memory (where the odd alignments of the two halves cost nothing)
// and copy it out with aligned 16-byte stores.  Dynamic shared memory: 32 cells x ldsh x 2 bytes (<= 37 KB).
__global__ void __launch_bounds__(256, 4) share_reduce_h16_kernel(const h16* __restrict__ x, int ld, int C,
                                                                  const TapRec* __restrict__ taps, int frame,
                                                                  h16* __restrict__ sh, int ldsh, int shift, int n_cells) {
  extern __shared__ uint4 share_rows[];
  const int grp = threadIdx.x >> 3, sub = threadIdx.x & 7;
  const int cell = blockIdx.x * 32 + grp;
  const bool live = cell < n_cells;
  const int pix = cell & (FEAT * FEAT - 1), chunk = cell >> 10;
  const int Cv = (C + 7) >> 3, S = shift + 2 * C;
  const uint64_t m1 = f2_pack(-1.f, -1.f);
  const float fr = (float)frame, inv_fr = 1.f / fr;
  const bool pow2 = (frame & (frame - 1)) == 0;            // then sum * (1 / frame) IS sum / frame
  unsigned short* row16 = reinterpret_cast<unsigned short*>(share_rows) + (size_t)grp * ldsh;
  h16* rowh = reinterpret_cast<h16*>(row16);
  if (live) {
    // padding of the row (never read as data): zeros, so that the buffer content is reproducible
    for (int k = sub; k < shift; k += 8) row16[k] = 0;
    for (int k = S + sub; k < ldsh; k += 8) row16[k] = 0;
    for (int cv = sub; cv < Cv; cv += 8) {
      const int c0 = 8 * cv;
      float mx[8];
      uint64_t sum[4];
#pragma unroll
      for (int k = 0; k < 8; ++k) mx[k] = -INFINITY;
#pragma unroll
      for (int k = 0; k < 4; ++k) sum[k] = 0ull;
#pragma unroll 2
      for (int f = 0; f < frame; ++f) {
        const int n = chunk * frame + f;
        const uint4 tp = __ldg(reinterpret_cast<const uint4*>(taps + ((size_t)n * (FEAT * FEAT) + pix) * 2));
        const h16* b = x + (size_t)n * (FEAT * FEAT) * ld + c0;
        const uint4 lt = __ldg(reinterpret_cast<const uint4*>(b + (size_t)(tp.x & 0xffffu) * ld));
        const uint4 rt = __ldg(reinterpret_cast<const uint4*>(b + (size_t)(tp.x >> 16) * ld));
        const uint4 lb = __ldg(reinterpret_cast<const uint4*>(b + (size_t)(tp.y & 0xffffu) * ld));
        const uint4 rb = __ldg(reinterpret_cast<const uint4*>(b + (size_t)(tp.y >> 16) * ld));
        const float o0f = __uint_as_float(tp.z), o1f = __uint_as_float(tp.w);
        const uint64_t o0 = f2_pack(o0f, o0f), o1 = f2_pack(o1f, o1f);
        const uint32_t a[4] = {lt.x, lt.y, lt.z, lt.w}, bq[4] = {rt.x, rt.y, rt.z, rt.w}, cq[4] = {lb.x, lb.y, lb.z, lb.w},
                       dq[4] = {rb.x, rb.y, rb.z, rb.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t v = warp_mix2(a[k], bq[k], cq[k], dq[k], o0, o1, m1);
          float v0, v1;
          f2_unpack(v, v0, v1);
          mx[2 * k] = fmaxf(mx[2 * k], v0);
          mx[2 * k + 1] = fmaxf(mx[2 * k + 1], v1);
          sum[k] = f2_add(sum[k], v);
        }
      }
      uint32_t om[4], oa[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float s0, s1;
        f2_unpack(sum[k], s0, s1);
        om[k] = pack_h16x2(mx[2 * k], mx[2 * k + 1]);
        oa[k] = pow2 ? pack_h16x2(s0 * inv_fr, s1 * inv_fr) : pack_h16x2(s0 / fr, s1 / fr);
      }
      if (c0 + 8 <= C) {
        store8_any(rowh + shift + c0, om);
        store8_any(rowh + shift + C + c0, oa);
      } else {
        // last, partial vector: the max half must not run into the mean half that follows it
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (c0 + k < C) {
            row16[shift + c0 + k] = (unsigned short)((om[k >> 1] >> (16 * (k & 1))) & 0xffffu);
            row16[shift + C + c0 + k] = (unsigned short)((oa[k >> 1] >> (16 * (k & 1))) & 0xffffu);
          }
        }
      }
    }
  }
  __syncwarp();
  if (live) {
    const uint4* src = reinterpret_cast<const uint4*>(row16);
    uint4* dst = reinterpret_cast<uint4*>(sh + (size_t)cell * ldsh);
    for (int v = sub; v < (ldsh >> 3); v += 8) dst[v] = src[v];
  }
}

// Second half (model_with_TSM.py:223-226): x[n][pix][coff + c] = warp_out(sh[chunk])[pix][shift + c], c < C2 = 2C.
__global__ void __launch_bounds__(256) share_out_h16_kernel(const h16* __restrict__ sh, int ldsh, int shift, int C2,
                                                            const TapRec* __restrict__ taps, int frame, h16* __restrict__ x,
                                                            int ld, int coff, int n_cells) {
  const int cell = blockIdx.x * 32 + (threadIdx.x >> 3), sub = threadIdx.x & 7;
  if (cell >= n_cells) return;
  const int n = cell >> 10;
  const int chunk = n / frame;
  const uint4 tp = __ldg(reinterpret_cast<const uint4*>(taps + (size_t)cell * 2 + 1));
  const float o0f = __uint_as_float(tp.z), o1f = __uint_as_float(tp.w);
  const uint64_t o0 = f2_pack(o0f, o0f), o1 = f2_pack(o1f, o1f), m1 = f2_pack(-1.f, -1.f);
  const h16* b = sh + (size_t)chunk * (FEAT * FEAT) * ldsh;
  const h16* p_lt = b + (size_t)(tp.x & 0xffffu) * ldsh;
  const h16* p_rt = b + (size_t)(tp.x >> 16) * ldsh;
  const h16* p_lb = b + (size_t)(tp.y & 0xffffu) * ldsh;
  const h16* p_rb = b + (size_t)(tp.y >> 16) * ldsh;
  h16* d = x + (size_t)cell * ld + coff - shift;                  // 16-byte aligned: ld % 8 == 0, (coff - shift) % 8 == 0
  unsigned short* d16 = reinterpret_cast<unsigned short*>(d);
  const int S = shift + C2, Nv = (S + 7) >> 3;
  for (int v = sub; v < Nv; v += 8) {
    const int s0 = 8 * v;
    const uint4 lt = __ldg(reinterpret_cast<const uint4*>(p_lt + s0));
    const uint4 rt = __ldg(reinterpret_cast<const uint4*>(p_rt + s0));
    const uint4 lb = __ldg(reinterpret_cast<const uint4*>(p_lb + s0));
    const uint4 rb = __ldg(reinterpret_cast<const uint4*>(p_rb + s0));
    const uint32_t a[4] = {lt.x, lt.y, lt.z, lt.w}, bq[4] = {rt.x, rt.y, rt.z, rt.w}, cq[4] = {lb.x, lb.y, lb.z, lb.w},
                   dq[4] = {rb.x, rb.y, rb.z, rb.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float v0, v1;
      f2_unpack(warp_mix2(a[k], bq[k], cq[k], dq[k], o0, o1, m1), v0, v1);
      o[k] = pack_h16x2(v0, v1);
    }
    if (s0 >= shift && s0 + 8 <= S) {
      *reinterpret_cast<uint4*>(d + s0) = make_uint4(o[0], o[1], o[2], o[3]);
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (s0 + k >= shift && s0 + k < S) d16[s0 + k] = (unsigned short)((o[k >> 1] >> (16 * (k & 1))) & 0xffffu);
    }
  }
}

// share == False: x_share = concat([x, x])  (model_with_TSM.py:227)
template <typename T>
__global__ void share_dup_kernel(T* __restrict__ x, int ld, int C, int coff, long long n_pix) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_pix * C) return;
  size_t pix = (size_t)(idx / C);
  int c = (int)(idx % C);
  T v = x[pix * ld + c];
  x[pix * ld + coff + c] = v;
  x[pix * ld + coff + C + c] = v;
}

// Pass-through channels of the ResBottleneck output: out[c] = leaky(x[c]) for c in [c0, c1)
// (model.py:109-113: y zero-extended, relu3(x + 0)).
template <typename T>
__global__ void res_tail_kernel(const T* __restrict__ x, int ldx, T* __restrict__ out, int ldo, int c0, int c1,
                                long long n_pix) {
  int span = c1 - c0;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_pix * span) return;
  size_t pix = (size_t)(idx / span);
  int c = c0 + (int)(idx % span);
  stf<T>(out, pix * ldo + c, leaky(ldf<T>(x, pix * ldx + c)));
}

// h16 product-path version: 8 channels (16 bytes) per thread; needs c0, c1, ldx, ldo multiples of 8.
__global__ void res_tail_vec_kernel(const h16* __restrict__ x, int ldx, h16* __restrict__ out, int ldo, int c0, int c1,
                                    long long n_pix) {
  const int span = (c1 - c0) >> 3;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_pix * span) return;
  const size_t pix = (size_t)(idx / span);
  const int c = c0 + 8 * (int)(idx % span);
  uint4 v = *reinterpret_cast<const uint4*>(x + pix * ldx + c);
  uint32_t* h = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = unpack_h16x2(h[i]);
    h[i] = pack_h16x2(leaky(f.x), leaky(f.y));
  }
  *reinterpret_cast<uint4*>(out + pix * ldo + c) = v;
}

// clr_conv2 (1x1 16->16 + BN + LeakyReLU), clr_conv3 (1x1 16->3) and the final
// dif = grey(con_rgb) - grey(inputs)   (model.py:268-269, 288).  w2[16][16] (in,out), w3[16][3].
template <typename T>
__global__ void clr_tail_kernel(const T* __restrict__ c16, int ld, const float* __restrict__ w2,
                                const float* __restrict__ b2, const float* __restrict__ w3,
                                const float* __restrict__ b3, const float* __restrict__ img,
                                float* __restrict__ rgb_out, float* __restrict__ dif_out, long long n_pix) {
  __shared__ float sw2[256], sb2[16], sw3[48], sb3[3];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) sw2[i] = w2[i];
  if (threadIdx.x < 16) sb2[threadIdx.x] = b2[threadIdx.x];
  if (threadIdx.x < 48) sw3[threadIdx.x] = w3[threadIdx.x];
  if (threadIdx.x < 3) sb3[threadIdx.x] = b3[threadIdx.x];
  __syncthreads();
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pix) return;
  float a[16], h[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) a[c] = ldf<T>(c16, (size_t)p * ld + c);
#pragma unroll
  for (int o = 0; o < 16; ++o) {
    float s = sb2[o];
#pragma unroll
    for (int c = 0; c < 16; ++c) s = fmaf(a[c], sw2[c * 16 + o], s);
    h[o] = leaky(s);
  }
  float rgb[3];
#pragma unroll
  for (int o = 0; o < 3; ++o) {
    float s = sb3[o];
#pragma unroll
    for (int c = 0; c < 16; ++c) s = fmaf(h[c], sw3[c * 3 + o], s);
    rgb[o] = s;
  }
  if (rgb_out) {
    rgb_out[3 * p] = rgb[0];
    rgb_out[3 * p + 1] = rgb[1];
    rgb_out[3 * p + 2] = rgb[2];
  }
  if (dif_out) {
    float g1 = rgb[0] * kGrayR + rgb[1] * kGrayG + rgb[2] * kGrayB;
    float g0 = img[3 * p] * kGrayR + img[3 * p + 1] * kGrayG + img[3 * p + 2] * kGrayB;
    dif_out[p] = g1 - g0;
  }
}

// tf.split of the dataset chunk (train_test_GSC.py:419, 806, 870, 900): [n,256,256,C] interleaved -> dense img / uv /
// reg / face planes.  One thread per (pixel, kept channel); o_* = first source channel of each piece, reg may be NULL.
__global__ void unpack_chunk_kernel(const float* __restrict__ chunk, int C, int o_uv, int o_reg, int o_face,
                                    float* __restrict__ img, float* __restrict__ uv, float* __restrict__ reg,
                                    float* __restrict__ face, long long n_pix) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_pix * 13) return;
  const long long p = idx / 13;
  const int k = (int)(idx - p * 13);
  const float* src = chunk + p * C;
  if (k < 3) img[p * 3 + k] = src[k];
  else if (k < 6) uv[p * 3 + (k - 3)] = src[o_uv + (k - 3)];
  else if (k < 12) { if (reg) reg[p * 6 + (k - 6)] = src[o_reg + (k - 6)]; }
  else face[p] = src[o_face];
}

// The same split, one block per 256 pixels: the block's slice of the chunk (256 x C contiguous floats) goes through shared
// memory with 16-byte loads (rows padded to an odd stride: the gathers below are conflict-free), and every output piece is
// written as one contiguous run.  The per-thread form above moves 4 bytes per thread behind a 64-bit division (measured
// 1.26 TB/s of DRAM traffic); dynamic shared memory = 256 x (C | 1) floats.
__global__ void __launch_bounds__(256) unpack_chunk_tile_kernel(const float* __restrict__ chunk, int C, int o_uv, int o_reg,
                                                                int o_face, float* __restrict__ img, float* __restrict__ uv,
                                                                float* __restrict__ reg, float* __restrict__ face,
                                                                long long n_pix) {
  extern __shared__ float unpack_tile[];
  const int Cp = C | 1, tid = threadIdx.x;
  const long long p0 = (long long)blockIdx.x * 256;
  const int np = n_pix - p0 < 256 ? (int)(n_pix - p0) : 256;
  const int nf = np * C;
  const float* src = chunk + p0 * C;
  if ((nf & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    for (int i = tid; i < (nf >> 2); i += 256) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(src) + i);
      const float e[4] = {v.x, v.y, v.z, v.w};
      int px = (4 * i) / C, k = 4 * i - px * C;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        unpack_tile[px * Cp + k] = e[q];
        if (++k == C) { k = 0; ++px; }
      }
    }
  } else {
    for (int i = tid; i < nf; i += 256) unpack_tile[(i / C) * Cp + (i % C)] = src[i];
  }
  __syncthreads();
  for (int j = tid; j < 3 * np; j += 256) {
    const int px = j / 3, k = j - 3 * px;
    img[p0 * 3 + j] = unpack_tile[px * Cp + k];
    uv[p0 * 3 + j] = unpack_tile[px * Cp + o_uv + k];
  }
  if (reg)
    for (int j = tid; j < 6 * np; j += 256) {
      const int px = j / 6, k = j - 6 * px;
      reg[p0 * 6 + j] = unpack_tile[px * Cp + o_reg + k];
    }
  for (int j = tid; j < np; j += 256) face[p0 + j] = unpack_tile[j * Cp + o_face];
}

// train_test_GSC.py:808-809: mask_pred = dif*face ; rgb = clip(rgb, 0, 1)
__global__ void caller_glue_kernel(const float* __restrict__ rgb, const float* __restrict__ dif,
                                   const float* __restrict__ face, float* __restrict__ rgb_c,
                                   float* __restrict__ mask_pred, long long n_pix) {
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pix) return;
  if (mask_pred) mask_pred[p] = dif[p] * face[p];
  if (rgb_c) {
#pragma unroll
    for (int c = 0; c < 3; ++c) rgb_c[3 * p + c] = fminf(fmaxf(rgb[3 * p + c], 0.f), 1.f);
  }
}

// train_test_GSC.py:711,718: out = clip(pred*m + inp*(1-m), 0, 1)
__global__ void composite_kernel(const float* __restrict__ pred, const float* __restrict__ inp,
                                 const float* __restrict__ m, float* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    float mm = m[i];
    out[i] = fminf(fmaxf(pred[i] * mm + inp[i] * (1.f - mm), 0.f), 1.f);
  }
}

// fp32 [n_pix][C] -> channels [0, C) of an activation buffer with row stride ld (bsr_share_layer entry)
template <typename T>
__global__ void pack_act_kernel(const float* __restrict__ in, int C, T* __restrict__ x, int ld, long long n_pix) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_pix * C) return;
  size_t pix = (size_t)(idx / C);
  int c = (int)(idx % C);
  stf<T>(x, pix * ld + c, in[idx]);
}

// debug: dense fp32 copy of channels [coff, coff+C) of an activation buffer
template <typename T>
__global__ void slice_to_f32_kernel(const T* __restrict__ x, int ld, int coff, int C, float* __restrict__ out,
                                    long long n_pix) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_pix * C) return;
  size_t pix = (size_t)(idx / C);
  int c = (int)(idx % C);
  out[idx] = ldf<T>(x, pix * ld + coff + c);
}

}  // namespace bsr
