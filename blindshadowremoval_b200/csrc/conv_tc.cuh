// Implicit-GEMM convolution on tcgen05 tensor cores (bf16 in, fp32 accumulate in TMEM).
//
// GEMM view:  D[128 pixels, BN couts] = sum over (tap, 64-channel block) A_tap[128, 64] . W_tap[BN, 64]^T
//   * A tile  = a BW x BH rectangle of output pixels (BW*BH = 128) of one image; for filter tap
//     (dy,dx) it is the same rectangle of the NHWC input shifted by (dy,dx), fetched by ONE 4-D TMA
//     box {64 ch, BW, BH, 1}: out-of-bounds rows/columns/channels are zero-filled by the TMA unit,
//     which is exactly TF 'SAME' padding (and the channel padding of 99/257/261-wide inputs).
//     Stride-2 convs use the same box with TMA element strides {1,2,2,1}.
//   * transposed 3x3/s2 convs (model.py:153) run as 4 sub-pixel phases (blockIdx.z): output
//     (2i+py, 2j+px) = sum over taps kh = py (mod 2), kw = px (mod 2) of in[i-(kh>>1), j-(kw>>1)] . W[kh,kw]
//     -> 4/2/2/1 taps, each phase a stride-1 gather GEMM with its own slice of K.
//   * W is packed once at load time: bf16 [cout_pad][taps * cin_pad64], K-major, fetched by 2-D TMA.
//   * 128-byte swizzle on both operands; UMMA M=128, N=BN (multiple of 16), K=16 per instruction.
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2-5 = epilogue (TMEM -> registers -> bias / residual adds / LeakyReLU -> global).
#pragma once
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "common.cuh"
#include "tc_common.cuh"

namespace bsr {

constexpr int TC_BM = 128;          // pixels per tile (UMMA M)
constexpr int TC_BK = 64;           // channels per k-block (128 B of bf16 = one swizzle row)
constexpr int TC_STAGES = 4;
constexpr int TC_THREADS = 192;
constexpr int TC_MAX_TAPS = 49;

struct TcWeights {
  bool ready = false;
  int kh = 0, kw = 0, cin = 0, cout = 0, transposed = 0;
  int row_packed = 0;     // conv1: input pre-packed as [N][H][W+8][8] bf16, one k-block = the 7(+1) pixels x 8 ch window
  int cin_pad = 0;        // multiple of 64
  int bn = 0;             // UMMA N per tile
  int n_tiles = 0;
  int taps = 0;
  bf16* dev = nullptr;    // [n_tiles*bn][taps*cin_pad]
  CUtensorMap map;        // 2-D {K_total, rows}, box {64, bn}
  int8_t tap_kh[TC_MAX_TAPS], tap_kw[TC_MAX_TAPS];   // packing order
  int phase_begin[5];
  void release() {
    if (dev) cudaFree(dev);
    dev = nullptr;
    ready = false;
  }
};

struct ConvTcParams {
  int n_img;
  int tiles_x, tiles_y;       // tiles per image in GEMM space
  int bw, bh;                 // tile rectangle
  int in_stride;              // 1 or 2: input coordinate = gemm coordinate * in_stride + tap offset
  int out_scale;              // 1 (conv) or 2 (transposed): output pixel = gemm pixel * out_scale + phase
  int OH, OW;
  int ncb;                    // 64-channel blocks per tap
  int bn;
  int phase_begin[5];
  int8_t dy[TC_MAX_TAPS], dx[TC_MAX_TAPS];
  int* errflag;
};

inline size_t conv_tc_smem_bytes(int bn) {
  return 1024 + (size_t)TC_STAGES * (TC_BM * 128 + (size_t)bn * 128) + 256;
}

// vectorised helpers for the epilogue ------------------------------------------------------
__device__ __forceinline__ void add_res16(const void* base, size_t pix, int ld, int c, int climit, float* v) {
  const bf16* p = (const bf16*)base + pix * ld + c;
  if (c + 16 <= climit) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      __nv_bfloat162 t = *reinterpret_cast<const __nv_bfloat162*>(&w[i]);
      v[2 * i] += __low2float(t);
      v[2 * i + 1] += __high2float(t);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (c + i < climit) v[i] += __bfloat162float(p[i]);
  }
}

__global__ void __launch_bounds__(TC_THREADS) conv_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                             const __grid_constant__ CUtensorMap tmB,
                                                             const ConvTcParams p, const EpiParams e) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_bytes = TC_BM * 128, b_bytes = (uint32_t)p.bn * 128;
  const uint32_t sA = smem_base;
  const uint32_t sB = sA + TC_STAGES * a_bytes;
  const uint32_t bars = sB + TC_STAGES * b_bytes;            // full[4], empty[4], tmem_full, tmem_ptr
  const uint32_t bar_full = bars, bar_empty = bars + 8 * TC_STAGES, bar_tmem = bars + 16 * TC_STAGES;
  const uint32_t tmem_slot = bar_tmem + 8;
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int n = blockIdx.x / tiles_per_img;
  const int tr = blockIdx.x % tiles_per_img;
  const int gy0 = (tr / p.tiles_x) * p.bh, gx0 = (tr % p.tiles_x) * p.bw;
  const int ntile = blockIdx.y, phase = blockIdx.z;
  const int t0 = p.phase_begin[phase], t1 = p.phase_begin[phase + 1];
  const int niter = (t1 - t0) * p.ncb;
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)p.bn) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_tmem, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < niter; ++it) {
        const int s = it % TC_STAGES;
        const uint32_t ph = (uint32_t)(it / TC_STAGES) & 1u;
        if (!mbar_wait(bar_empty + 8 * s, ph ^ 1u, p.errflag, 1)) break;
        const int t = t0 + it / p.ncb, cb = it % p.ncb;
        mbar_expect_tx(bar_full + 8 * s, a_bytes + b_bytes);
        tma_load_4d(sA + s * a_bytes, &tmA, bar_full + 8 * s, cb * TC_BK, gx0 * p.in_stride + p.dx[t],
                    gy0 * p.in_stride + p.dy[t], n);
        tma_load_2d(sB + s * b_bytes, &tmB, bar_full + 8 * s, (t * p.ncb + cb) * TC_BK, ntile * p.bn);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(TC_BM, p.bn);
      bool ok = true;
      for (int it = 0; it < niter && ok; ++it) {
        const int s = it % TC_STAGES;
        const uint32_t ph = (uint32_t)(it / TC_STAGES) & 1u;
        ok = mbar_wait(bar_full + 8 * s, ph, p.errflag, 2);
        if (!ok) break;
        tc_fence_after();
        const uint64_t da = umma_desc_sw128(sA + s * a_bytes), db = umma_desc_sw128(sB + s * b_bytes);
#pragma unroll
        for (int k = 0; k < TC_BK / 16; ++k)
          umma_bf16(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (it | k) != 0 ? 1u : 0u);
        umma_commit(bar_empty + 8 * s);
      }
      umma_commit(bar_tmem);
    }
  } else {
    // ---- epilogue: warp w reads TMEM lanes 32*(w%4) .. +31 (row = pixel of the tile)
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int gy = gy0 + r / p.bw, gx = gx0 + r % p.bw;
    const int oy = gy * p.out_scale + (phase >> 1), ox = gx * p.out_scale + (phase & 1);
    const size_t pix = ((size_t)n * p.OH + oy) * p.OW + ox;
    const bool ok = mbar_wait(bar_tmem, 0, p.errflag, 3);
    tc_fence_after();
    if (ok) {
      const bool vec_ok = (e.out_ld % 8 == 0) && (e.out_coff % 8 == 0);
      const int s_img = gy * p.OW * p.out_scale + gx;      // OUT_QKV only (out_scale 1): pixel index inside the image
      for (int j = 0; j < p.bn; j += 16) {
        const int c = ntile * p.bn + j;
        if (c >= e.out_c) break;               // warp-uniform
        float v[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)j, v);
        if (c + 16 <= e.out_c && (vec_ok || e.out_mode != OUT_T)) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += __ldg(e.bias + c + i);      // bias is zero-padded past cout
          if (e.res1 != nullptr && c < e.res1_c) add_res16(e.res1, pix, e.res1_ld, c, e.res1_c, v);
          if (e.res2 != nullptr && c < e.res2_c) add_res16(e.res2, pix, e.res2_ld, c, e.res2_c, v);
          if (e.act) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = leaky(v[i]);
          }
          if (e.out_mode == OUT_F32) {
            float* dst = (float*)e.out + pix * e.out_ld + e.out_coff + c;
#pragma unroll
            for (int i = 0; i < 16; ++i) dst[i] = v[i];
          } else if (e.out_mode == OUT_QKV && c >= 256) {
            // g -> V^T[n][c-256][s]: for a fixed channel the 32 lanes of a warp write consecutive tokens
            bf16* dst = (bf16*)e.out2 + ((size_t)n * 128 + (c - 256)) * e.spatial + s_img;
#pragma unroll
            for (int i = 0; i < 16; ++i) dst[(size_t)i * e.spatial] = __float2bfloat16_rn(v[i]);
          } else {
            uint4 o0, o1;
            o0.x = pack_bf16x2(v[0], v[1]); o0.y = pack_bf16x2(v[2], v[3]);
            o0.z = pack_bf16x2(v[4], v[5]); o0.w = pack_bf16x2(v[6], v[7]);
            o1.x = pack_bf16x2(v[8], v[9]); o1.y = pack_bf16x2(v[10], v[11]);
            o1.z = pack_bf16x2(v[12], v[13]); o1.w = pack_bf16x2(v[14], v[15]);
            const size_t off = e.out_mode == OUT_QKV ? pix * 256 + c : pix * e.out_ld + e.out_coff + c;
            uint4* dst = reinterpret_cast<uint4*>((bf16*)e.out + off);
            dst[0] = o0;
            dst[1] = o1;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) epi_store<bf16>(e, pix, c + i, v[i]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------
// host side
inline int configure_tc_kernels_conv() {
  cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)conv_tc_smem_bytes(256));
  return e == cudaSuccess ? 0 : (int)e;
}

inline uint16_t f32_to_bf16_bits(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

// Pack canonical fp32 [tap][cin][cout] into bf16 [cout_pad][taps*cin_pad] (K-major).  Returns false
// with empty *why when the layer is simply not eligible (kept on the CUDA-core kernel).
inline bool pack_tc_weights(TmaEncoder& tma, const std::string& name, int kh, int kw, int cin, int cout, int transposed,
                            const std::vector<float>& w, TcWeights* out, std::string* why) {
  why->clear();
  if (name == "clr_conv2" || name == "clr_conv3") return false;   // fused into the colour tail kernel
  if (const char* dis = getenv("BSR_TC_DISABLE")) {
    std::string d = std::string(",") + dis + ",";
    if (d.find("," + name + ",") != std::string::npos) return false;
    std::string base = name.substr(0, name.find('.') == std::string::npos ? name.size() : name.find('.'));
    if (name.find('.') != std::string::npos && d.find(",res." + name.substr(name.find('.') + 1) + ",") != std::string::npos)
      return false;
    (void)base;
  }
  TcWeights& t = *out;
  t.kh = kh; t.kw = kw; t.cin = cin; t.cout = cout; t.transposed = transposed;
  if (name == "conv1") {
    // 7x7 conv over 3 channels (model.py:203): K per filter ROW = 7 taps x 8 (3 real + 5 zero) channels + one
    // zero-weight pixel = 64 = one 128-byte swizzle row, fetched as an overlapping window of the packed image.
    if (kh != 7 || kw != 7 || cin > 8 || transposed) { *why = "conv1 must be 7x7 with <= 8 input channels"; return false; }
    t.row_packed = 1; t.cin_pad = 64; t.taps = 7; t.bn = (cout + 15) / 16 * 16; t.n_tiles = 1;
    for (int a = 0; a < 7; ++a) { t.tap_kh[a] = (int8_t)a; t.tap_kw[a] = 3; }
    t.phase_begin[0] = 0;
    for (int i = 1; i < 5; ++i) t.phase_begin[i] = 7;
    const size_t K = 7 * 64, rows = t.bn;
    std::vector<uint16_t> host(rows * K, 0);
    for (int a = 0; a < 7; ++a)
      for (int b = 0; b < 7; ++b)
        for (int c = 0; c < cin; ++c)
          for (int o = 0; o < cout; ++o)
            host[(size_t)o * K + a * 64 + b * 8 + c] = f32_to_bf16_bits(w[((size_t)(a * 7 + b) * cin + c) * cout + o]);
    if (cudaMalloc(&t.dev, host.size() * 2) != cudaSuccess) { *why = "cudaMalloc failed"; return false; }
    if (cudaMemcpy(t.dev, host.data(), host.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess) { *why = "cudaMemcpy failed"; return false; }
    uint64_t dims[2] = {K, rows}, strides[1] = {K * 2};
    uint32_t box[2] = {TC_BK, (uint32_t)t.bn};
    if (!tma.encode_bf16(&t.map, t.dev, 2, dims, strides, box, nullptr)) { *why = tma.last_error; return false; }
    t.ready = true;
    return true;
  }
  t.cin_pad = (cin + 63) / 64 * 64;
  t.taps = kh * kw;
  if (t.taps > TC_MAX_TAPS) return false;
  if (cout == 384) { t.bn = 128; t.n_tiles = 3; }
  else if (cout <= 256) { t.bn = (cout + 15) / 16 * 16; t.n_tiles = 1; }
  else { t.n_tiles = (cout + 255) / 256; t.bn = ((cout + t.n_tiles - 1) / t.n_tiles + 15) / 16 * 16; }
  // tap order: phase-major for transposed convs
  int nt = 0;
  if (!transposed) {
    for (int a = 0; a < kh; ++a)
      for (int b = 0; b < kw; ++b) { t.tap_kh[nt] = (int8_t)a; t.tap_kw[nt] = (int8_t)b; ++nt; }
    t.phase_begin[0] = 0;
    for (int i = 1; i < 5; ++i) t.phase_begin[i] = nt;
  } else {
    if (kh != 3 || kw != 3) { *why = "transposed conv must be 3x3"; return false; }
    for (int ph = 0; ph < 4; ++ph) {
      t.phase_begin[ph] = nt;
      int py = ph >> 1, px = ph & 1;
      for (int a = py; a < 3; a += 2)
        for (int b = px; b < 3; b += 2) { t.tap_kh[nt] = (int8_t)a; t.tap_kw[nt] = (int8_t)b; ++nt; }
    }
    t.phase_begin[4] = nt;
  }
  const size_t K = (size_t)t.taps * t.cin_pad, rows = (size_t)t.n_tiles * t.bn;
  std::vector<uint16_t> host(rows * K, 0);
  for (int ti = 0; ti < t.taps; ++ti) {
    const int src_tap = t.tap_kh[ti] * kw + t.tap_kw[ti];
    for (int c = 0; c < cin; ++c)
      for (int o = 0; o < cout; ++o)
        host[(size_t)o * K + (size_t)ti * t.cin_pad + c] = f32_to_bf16_bits(w[((size_t)src_tap * cin + c) * cout + o]);
  }
  if (cudaMalloc(&t.dev, host.size() * 2) != cudaSuccess) { *why = "cudaMalloc failed"; return false; }
  if (cudaMemcpy(t.dev, host.data(), host.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess) { *why = "cudaMemcpy failed"; return false; }
  uint64_t dims[2] = {K, rows}, strides[1] = {K * 2};
  uint32_t box[2] = {TC_BK, (uint32_t)t.bn};
  if (!tma.encode_bf16(&t.map, t.dev, 2, dims, strides, box, nullptr)) { *why = tma.last_error; return false; }
  t.ready = true;
  return true;
}

struct TmapKey {
  const void* p; int ld, coff, cin, H, W, n, stride, bw, bh;
  bool operator<(const TmapKey& o) const {
    return std::tie(p, ld, coff, cin, H, W, n, stride, bw, bh) <
           std::tie(o.p, o.ld, o.coff, o.cin, o.H, o.W, o.n, o.stride, o.bw, o.bh);
  }
};

inline int launch_conv_tc(TmaEncoder& tma, const TcWeights& t, const void* in, int in_ld, int in_coff, int H, int W,
                          int stride, int n, const EpiParams& e, int* errflag, cudaStream_t st, int* launches) {
  static thread_local std::map<TmapKey, CUtensorMap> cache;
  ConvTcParams p;
  memset(&p, 0, sizeof p);
  int GH, GW;
  if (t.transposed) {
    GH = H; GW = W; p.in_stride = 1; p.out_scale = 2; p.OH = 2 * H; p.OW = 2 * W;
    for (int i = 0; i < t.taps; ++i) { p.dy[i] = (int8_t)(-(t.tap_kh[i] >> 1)); p.dx[i] = (int8_t)(-(t.tap_kw[i] >> 1)); }
  } else {
    p.in_stride = stride; p.out_scale = 1;
    p.OH = (H + stride - 1) / stride; p.OW = (W + stride - 1) / stride;
    GH = p.OH; GW = p.OW;
    int tot_h = (p.OH - 1) * stride + t.kh - H; if (tot_h < 0) tot_h = 0;
    int tot_w = (p.OW - 1) * stride + t.kw - W; if (tot_w < 0) tot_w = 0;
    for (int i = 0; i < t.taps; ++i) { p.dy[i] = (int8_t)(t.tap_kh[i] - tot_h / 2); p.dx[i] = (int8_t)(t.tap_kw[i] - tot_w / 2); }
  }
  p.bw = GW < TC_BM ? GW : TC_BM;
  p.bh = TC_BM / p.bw;
  if (GW % p.bw || GH % p.bh) { tma.last_error = "spatial size not tileable into 128-pixel rectangles"; return -1; }
  p.tiles_x = GW / p.bw; p.tiles_y = GH / p.bh;
  p.n_img = n; p.ncb = t.cin_pad / TC_BK; p.bn = t.bn;
  for (int i = 0; i < 5; ++i) p.phase_begin[i] = t.phase_begin[i];
  p.errflag = errflag;
  if ((in_ld % 8) || (in_coff % 8)) { tma.last_error = "input channel stride/offset must be multiples of 8"; return -2; }
  if (t.row_packed) { for (int i = 0; i < t.taps; ++i) p.dx[i] = 0; }
  TmapKey key{in, in_ld, in_coff, t.cin, H, W, n, p.in_stride, p.bw, p.bh};
  auto it = cache.find(key);
  if (it == cache.end()) {
    CUtensorMap m;
    uint64_t dims[4] = {(uint64_t)t.cin, (uint64_t)W, (uint64_t)H, (uint64_t)n};
    uint64_t strides[3] = {(uint64_t)in_ld * 2, (uint64_t)W * in_ld * 2, (uint64_t)H * W * in_ld * 2};
    if (t.row_packed) {
      // packed image [N][H][W + 8][8]: window of output column x = 64 contiguous elements starting at packed
      // pixel x (= image pixel x-3); consecutive windows overlap (dim-1 stride 16 B < dim-0 extent 128 B)
      dims[0] = 64;
      strides[0] = 16; strides[1] = (uint64_t)(W + 8) * 16; strides[2] = (uint64_t)H * (W + 8) * 16;
    }
    uint32_t box[4] = {TC_BK, (uint32_t)(p.bw * p.in_stride), (uint32_t)(p.bh * p.in_stride), 1};
    uint32_t es[4] = {1, (uint32_t)p.in_stride, (uint32_t)p.in_stride, 1};
    if (!tma.encode_bf16(&m, (void*)((const bf16*)in + in_coff), 4, dims, strides, box, es)) return -3;
    if (cache.size() > 4096) cache.clear();
    it = cache.emplace(key, m).first;
  }
  dim3 grid((unsigned)(n * p.tiles_x * p.tiles_y), (unsigned)t.n_tiles, t.transposed ? 4u : 1u);
  conv_tc_kernel<<<grid, TC_THREADS, conv_tc_smem_bytes(t.bn), st>>>(it->second, t.map, p, e);
  (*launches)++;
  return 0;
}

}  // namespace bsr
