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

// 16-bit activation / weight storage of the tensor-core path.  Default: IEEE binary16 (11 significand bits; the
// kind::f16 UMMA takes it at the same rate as bfloat16 and the end-to-end error drops ~8x, which is what north_star's
// 1e-2 max-abs needs).  Conversions saturate to +-65504 instead of producing inf.  -DBSR_ACT_BF16 builds the bfloat16
// variant (A/B error measurements only).
#ifdef BSR_ACT_BF16
typedef __nv_bfloat16 h16;
typedef __nv_bfloat162 h16x2;
#define BSR_ACT_DTYPE_NAME "bf16"
constexpr uint32_t kUmmaOperandFmt = 1u;      // instruction-descriptor A/B format field: 1 = bf16
__device__ __forceinline__ float h16_to_f32(h16 v) { return __bfloat162float(v); }
__device__ __forceinline__ h16 f32_to_h16(float v) { return __float2bfloat16_rn(v); }
__device__ __forceinline__ uint32_t pack_h16x2(float lo, float hi) {
  h16x2 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float2 unpack_h16x2(uint32_t w) { return __bfloat1622float2(*reinterpret_cast<const h16x2*>(&w)); }
#else
typedef __half h16;
typedef __half2 h16x2;
#define BSR_ACT_DTYPE_NAME "f16"
constexpr uint32_t kUmmaOperandFmt = 0u;      // 0 = f16
__device__ __forceinline__ float h16_to_f32(h16 v) { return __half2float(v); }
__device__ __forceinline__ uint32_t pack_h16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));     // first source -> upper half
  return r;
}
__device__ __forceinline__ h16 f32_to_h16(float v) {
  uint16_t r;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(r) : "f"(v));
  return __ushort_as_half(r);
}
__device__ __forceinline__ float2 unpack_h16x2(uint32_t w) { return __half22float2(*reinterpret_cast<const h16x2*>(&w)); }
#endif

template <typename T> __device__ __forceinline__ float ldf(const T* p, size_t i);
template <> __device__ __forceinline__ float ldf<float>(const float* p, size_t i) { return p[i]; }
template <> __device__ __forceinline__ float ldf<h16>(const h16* p, size_t i) { return h16_to_f32(p[i]); }
template <typename T> __device__ __forceinline__ void stf(T* p, size_t i, float v);
template <> __device__ __forceinline__ void stf<float>(float* p, size_t i, float v) { p[i] = v; }
template <> __device__ __forceinline__ void stf<h16>(h16* p, size_t i, float v) { p[i] = f32_to_h16(v); }

__device__ __forceinline__ float leaky(float v) { return v >= 0.f ? v : kLeaky * v; }

// 256-bit global store / load (sm_100: STG.256 / LDG.256): a thread that owns 32 contiguous, 32-byte aligned bytes moves
// them with ONE instruction - half the L1 wavefronts of two 128-bit accesses when every lane touches a different line.
__device__ __forceinline__ void st_global_256(void* p, const uint4& a, const uint4& b) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w),
               "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}
__device__ __forceinline__ void ld_global_256(const void* p, uint4& a, uint4& b) {
  asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w),
               "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p) : "memory");
}

// Packed fp32 FMA of sm_100 (fma.rn.f32x2 -> FFMA2): two IEEE fp32 FMAs per issued instruction, bit-identical to two
// fmaf().  (d0, d1) += a * (w0, w1)   and   (d0, d1) += (a0, a1) * (w0, w1).  The register-pair moves fold away.
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a, float w0, float w1) {
  uint64_t d, aa, ww;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(d0), "f"(d1));
  asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
  asm("mov.b64 %0, {%1, %2};" : "=l"(ww) : "f"(w0), "f"(w1));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(aa), "l"(ww));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
}
// (r0, r1) = (a0, a1) * b + c   and   (d0, d1) += (a0, a1)
__device__ __forceinline__ void fma2_bc(float& r0, float& r1, float a0, float a1, float b, float c) {
  uint64_t d, aa, bb, cc;
  asm("mov.b64 %0, {%1, %2};" : "=l"(aa) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
  asm("mov.b64 %0, {%1, %1};" : "=l"(cc) : "f"(c));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(aa), "l"(bb), "l"(cc));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r0), "=f"(r1) : "l"(d));
}
__device__ __forceinline__ void fadd2(float& d0, float& d1, float a0, float a1) {
  uint64_t d, aa;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(d0), "f"(d1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(aa) : "f"(a0), "f"(a1));
  asm("add.rn.f32x2 %0, %0, %1;" : "+l"(d) : "l"(aa));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
}
// (r0, r1) = (a0, a1) * b
__device__ __forceinline__ void fmul2_b(float& r0, float& r1, float a0, float a1, float b) {
  uint64_t d, aa, bb;
  asm("mov.b64 %0, {%1, %2};" : "=l"(aa) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(aa), "l"(bb));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r0), "=f"(r1) : "l"(d));
}
__device__ __forceinline__ void ffma2v(float& d0, float& d1, float a0, float a1, float w0, float w1) {
  uint64_t d, aa, ww;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(d0), "f"(d1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(aa) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(ww) : "f"(w0), "f"(w1));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(aa), "l"(ww));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
}

// Experiment / bring-up switches, read from the environment ONCE per handle (bsr_create), never on the launch path.
struct Knobs {
  int ablate = 0;          // BSR_ABLATE (profiling only): 1 = no epilogue stores, 2 = no MMA, 4 = no A-tile TMA, 8 = role timers
  int no_pdl = 0;          // BSR_NO_PDL: launch without programmatic dependent launch
  int no_tma_store = 0;    // BSR_NO_TMA_STORE: per-thread global stores instead of the staged TMA-store epilogues
  int st_bufs = 0;         // BSR_ST_BUFS: force the number of staging buffers (1 / 2) where it fits
  int no_fuse_w = 0;       // BSR_NO_FUSE_W: NonLocal output conv as its own launch
  int host_chunk = 0;      // BSR_HOST_CHUNK: images per pipelined host-path chunk
  int host_full_uv = 0;    // BSR_HOST_FULL_UV: the fp32 host path uploads uv / reg in full instead of the rows the model reads
  int no_graph = 0;        // BSR_NO_GRAPH: never replay micro-batches from captured CUDA graphs
  int no_halo = 0;         // BSR_NO_HALO: fused transposed convs fetch every shifted A tile separately (round-1 behaviour)
  int no_hole_inplace = 0; // BSR_NO_HOLE_INPLACE: the hole mask copies into the other buffer even when the strides agree
  int share_v1 = 0;        // BSR_SHARE_V1: warp-per-cell ShareLayer kernels on the 16-bit path too (A/B measurements)
  int unpack_v1 = 0;       // BSR_UNPACK_V1: per-thread chunk split instead of the shared-memory tile form (A/B measurements)
  int no_halo3 = 0;        // BSR_NO_HALO3: res conv2 on the generic kernel (nine shifted A tiles) instead of conv3x3_halo.cuh
};
// Launch-plan counters of one forward (bsr_plan_counter).
struct PlanCounters { int resident = 0, pinned = 0, staged = 0, attn_fused = 0, graph_replays = 0, halo3 = 0; };

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
  void* out2;        // OUT_QKV: base of g: V^T[n][128][spatial] (v_natural = 0) or V[pix][128] (v_natural = 1)
  int spatial;       // OUT_QKV: pixels per image (1024)
  int v_natural;
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
      if (e.v_natural) {
        stf<T>((T*)e.out2, pix * 128 + (c - 256), v);
      } else {
        size_t n = pix / e.spatial, s = pix % e.spatial;
        stf<T>((T*)e.out2, (n * 128 + (c - 256)) * e.spatial + s, v);
      }
    }
  }
}

}  // namespace bsr
