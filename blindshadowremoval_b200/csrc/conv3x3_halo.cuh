// 3x3 / stride 1 convolution, 128 -> 128 channels (the middle conv of every ResBottleneck, model.py:17-19), as an implicit
// GEMM whose A operand is ONE halo tile per K block instead of nine shifted copies of the pixel tile.
//
// The generic kernel (conv_tc.cuh) fetches a 128-pixel A tile per (tap, K block): 9 x 2 x 16 KB of activations plus
// 9 x 2 x 16 KB of streamed weights per 128 pixels = 583 KB through the TMA unit, which delivers about 64 B per cycle and
// SM - the layer was bound by that, not by its MMAs.  Here a CTA works on a 16 x 16 pixel region:
//   * A: for each of the two 64-channel K blocks ONE TMA box of 18 x 18 pixels (the region plus its 3x3 halo, zero-filled
//     outside the image) = 324 rows of 128 bytes, row index = y' * 18 + x'.  The region is two 16 x 8 sub-tiles (left and
//     right); the A operand of tap (dy, dx) for a sub-tile is a VIEW of the halo tile: it starts at row
//     (dy + 1) * 18 + (dx + 1) + 8 * half and its sixteen 8-row groups (one image row of the sub-tile each) are
//     18 rows = 2304 bytes apart, which is exactly what the descriptor's stride-byte-offset field expresses.  The
//     128-byte swizzle is a function of the shared-memory address bits, so a group may start on any 128-byte row of a
//     TMA-written tile (tools/halo_probe.py; the up3 halo tiles rely on the same fact).
//   * B: the (tap, K block) weight blocks (128 rows x 128 bytes) stream through a small ring and each one feeds BOTH
//     sub-tiles, so the weight traffic per pixel halves.
// Per 256 pixels: 2 x 41 KB + 18 x 16 KB = 378 KB instead of 1166 KB.
// TMEM: four 128-column accumulators; region t uses the pair 2 * (t & 1), so the epilogue of one region overlaps the MMAs
// of the next.  Warp roles as in conv_tc.cuh: warps 0-15 epilogue, 16 TMA producer, 17 MMA issuer.
#pragma once
#include "conv_tc.cuh"

namespace bsr {

constexpr int H3_PITCH = 18;                          // halo tile: 18 x 18 pixels
constexpr int H3_ROWS = H3_PITCH * H3_PITCH;          // 324 rows of 128 bytes
constexpr int H3_A_BYTES = 42 * 1024;                 // one K block of the halo tile, padded to a 1024-byte multiple
constexpr int H3_A_BOX_BYTES = H3_ROWS * 128;         // bytes the TMA box delivers
constexpr int H3_B_BYTES = 128 * 128;                 // one (tap, K block) weight block
constexpr int H3_B_STAGES = 3;
constexpr int H3_SMEM = 1024 + 4 * H3_A_BYTES + H3_B_STAGES * H3_B_BYTES + 1024;
// A-operand descriptor, high word: stride between 8-row groups = one halo row of 18 pixels
constexpr uint32_t kH3DescHiA = ((uint32_t)(H3_PITCH * 128) >> 4) | (1u << 14) | (2u << 29);

struct Halo3Params {
  int n_img, H, W;              // image count and size (H, W multiples of 16)
  int tiles_x, tiles_y, total_tiles;
  const float* bias;
  int act;
  void* out; int out_ld, out_coff;
  int ablate;
  int* errflag;
};

__device__ __forceinline__ void umma_h16_lo_hi(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t idesc,
                                                uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %6};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kUmmaDescHi), "r"(a_hi)
      : "memory");
}

__global__ void __launch_bounds__(TC_THREADS, 1) conv3x3_halo_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                     const __grid_constant__ CUtensorMap tmB,
                                                                     const __grid_constant__ Halo3Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  // barriers, TMEM slot and bias FIRST (1 KB), the TMA-written tiles behind them
  const uint32_t bars = smem_base;
  const uint32_t sA = smem_base + 1024u;                                // [buffer 2][K block 2] halo tiles
  const uint32_t sB = sA + 4u * H3_A_BYTES;                             // weight ring
  const uint32_t bar_afull = bars, bar_aempty = bars + 16, bar_bfull = bars + 32, bar_bempty = bars + 64;
  const uint32_t bar_tfull = bars + 96, bar_tempty = bars + 112, tmem_slot = bars + 128;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_al + (tmem_slot - smem_base));
  float* bias_s = reinterpret_cast<float*>(smem_al + (bars + 256 - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = p.total_tiles, tiles_per_img = p.tiles_x * p.tiles_y;
  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_afull + 8 * i, 1);
      mbar_init(bar_aempty + 8 * i, 1);
      mbar_init(bar_tfull + 8 * i, 1);
      mbar_init(bar_tempty + 8 * i, TC_EPI_WARPS);
    }
    for (int i = 0; i < H3_B_STAGES; ++i) {
      mbar_init(bar_bfull + 8 * i, 1);
      mbar_init(bar_bempty + 8 * i, 1);
    }
    fence_barrier_init();
  }
  if (threadIdx.x < 128) bias_s[threadIdx.x] = __ldg(p.bias + threadIdx.x);
  if (warp == TC_EPI_WARPS + 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_launch_dependents();

  if (warp == TC_EPI_WARPS) {
    // ================= TMA producer =================
    const bool leader = elect_one();
    bool ok = true;
    uint32_t bs = 0, bph = 0, lt = 0;
    pdl_wait();
    auto load_a = [&](int tile, uint32_t lti) {          // both K blocks of one region's halo tile
      const uint32_t buf = lti & 1u, par = (lti >> 1) & 1u;
      ok = mbar_wait(bar_aempty + 8 * buf, par ^ 1u, p.errflag, 1);
      if (!ok) return;
      if (leader) {
        const int n = tile / tiles_per_img, tr = tile % tiles_per_img;
        const int x0 = (tr % p.tiles_x) * 16 - 1, y0 = (tr / p.tiles_x) * 16 - 1;
        if (!(p.ablate & 4)) {
          mbar_expect_tx(bar_afull + 8 * buf, 2u * H3_A_BOX_BYTES);
          tma_load_4d(sA + (buf * 2u) * H3_A_BYTES, &tmA, bar_afull + 8 * buf, 0, x0, y0, n);
          tma_load_4d(sA + (buf * 2u + 1u) * H3_A_BYTES, &tmA, bar_afull + 8 * buf, 64, x0, y0, n);
        } else {
          mbar_arrive(bar_afull + 8 * buf);
        }
      }
      __syncwarp();
    };
    int tile = blockIdx.x;
    if (tile < total_tiles) load_a(tile, 0);
    for (; tile < total_tiles && ok; tile += gridDim.x, ++lt) {
      for (int j = 0; j < 18 && ok; ++j) {             // j = kb * 9 + tap: weight K coordinate (tap * 2 + kb) * 64
        const int kb = j / 9, tap = j - kb * 9;
        ok = mbar_wait(bar_bempty + 8 * bs, bph ^ 1u, p.errflag, 1);
        if (!ok) break;
        if (leader) {
          mbar_expect_tx(bar_bfull + 8 * bs, H3_B_BYTES);
          tma_load_2d(sB + bs * H3_B_BYTES, &tmB, bar_bfull + 8 * bs, (tap * 2 + kb) * TC_BK, 0);
        }
        __syncwarp();
        if (++bs == H3_B_STAGES) { bs = 0; bph ^= 1u; }
        // the next region's halo tile goes out once this region's first K block is on its way: by then the MMA warp
        // has long left the buffer it overwrites (it is consuming this region's weights)
        if (j == 8 && tile + (int)gridDim.x < total_tiles) load_a(tile + gridDim.x, lt + 1);
      }
    }
  } else if (warp == TC_EPI_WARPS + 1) {
    // ================= MMA issuer =================
    const bool leader = elect_one();
    bool ok = true;
    uint32_t bs = 0, bph = 0, lt = 0;
    const uint32_t idesc = umma_idesc_h16(TC_BM, 128);
    for (int tile = blockIdx.x; tile < total_tiles && ok; tile += gridDim.x, ++lt) {
      const uint32_t buf = lt & 1u, par = (lt >> 1) & 1u;
      ok = mbar_wait(bar_tempty + 8 * buf, par ^ 1u, p.errflag, 4, true);
      if (!ok) break;
      ok = mbar_wait(bar_afull + 8 * buf, par, p.errflag, 2, true);
      if (!ok) break;
      tc_fence_after();
      const uint32_t acc = tmem_base + buf * 256u;
      for (int j = 0; j < 18 && ok; ++j) {
        const int kb = j / 9, tap = j - kb * 9;
        ok = mbar_wait(bar_bfull + 8 * bs, bph, p.errflag, 2, true);
        if (!ok) break;
        tc_fence_after();
        if (leader && !(p.ablate & 2)) {
          const int dy = tap / 3, dx = tap - dy * 3;           // = (dy + 1), (dx + 1) of the centred tap offsets
          const uint32_t a_lo0 = umma_desc_lo(sA + (buf * 2u + (uint32_t)kb) * H3_A_BYTES) + (uint32_t)(dy * H3_PITCH + dx) * 8u;
          const uint32_t b_lo = umma_desc_lo(sB + bs * H3_B_BYTES);
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const uint32_t a_lo = a_lo0 + (uint32_t)half * 64u, d = acc + (uint32_t)half * 128u;
            umma_h16_lo_hi(d, a_lo, kH3DescHiA, b_lo, idesc, j == 0 ? 0u : 1u);
            umma_h16_lo_hi(d, a_lo + 2, kH3DescHiA, b_lo + 2, idesc, 1u);
            umma_h16_lo_hi(d, a_lo + 4, kH3DescHiA, b_lo + 4, idesc, 1u);
            umma_h16_lo_hi(d, a_lo + 6, kH3DescHiA, b_lo + 6, idesc, 1u);
          }
        }
        if (leader) umma_commit(bar_bempty + 8 * bs);
        __syncwarp();
        if (++bs == H3_B_STAGES) { bs = 0; bph ^= 1u; }
      }
      if (leader) {
        umma_commit(bar_aempty + 8 * buf);
        umma_commit(bar_tfull + 8 * buf);
      }
      __syncwarp();
    }
  } else {
    // ================= epilogue: warp = (lane quarter q, sub-tile half, 64-column half) =================
    const int q = warp & 3, cg = warp >> 2, half = cg >> 1, chalf = cg & 1;
    const int r = q * 32 + lane;                          // TMEM lane = pixel of the sub-tile: row r / 8, column r % 8
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    bool ok = true;
    uint32_t lt = 0;
    pdl_wait();
    for (int tile = blockIdx.x; tile < total_tiles && ok; tile += gridDim.x, ++lt) {
      const uint32_t buf = lt & 1u, par = (lt >> 1) & 1u;
      const int n = tile / tiles_per_img, tr = tile % tiles_per_img;
      const int y = (tr / p.tiles_x) * 16 + (r >> 3), x = (tr % p.tiles_x) * 16 + half * 8 + (r & 7);
      ok = mbar_wait(bar_tfull + 8 * buf, par, p.errflag, 3);
      if (!ok) break;
      tc_fence_after();
      const uint32_t ta = tmem_base + lane_addr + buf * 256u + (uint32_t)half * 128u + (uint32_t)chalf * 64u;
      float v[4][16];
#pragma unroll
      for (int k = 0; k < 4; ++k) tmem_ld16_nowait(ta + 16u * k, v[k]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
      h16* const dst = (h16*)p.out + (((size_t)n * p.H + y) * p.W + x) * p.out_ld + p.out_coff + chalf * 64;
      uint32_t o[4][8];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(bias_s + chalf * 64 + 16 * k + i);
          float a0 = v[k][i] + b4.x, a1 = v[k][i + 1] + b4.y, a2 = v[k][i + 2] + b4.z, a3 = v[k][i + 3] + b4.w;
          if (p.act) {
            a0 = fmaxf(a0, kLeaky * a0); a1 = fmaxf(a1, kLeaky * a1);
            a2 = fmaxf(a2, kLeaky * a2); a3 = fmaxf(a3, kLeaky * a3);
          }
          o[k][i >> 1] = pack_h16x2(a0, a1);
          o[k][(i >> 1) + 1] = pack_h16x2(a2, a3);
        }
      }
      // lane L stores chunk (j + L) % 4 in store j (see the fused transposed-conv epilogue in conv_tc.cuh)
      const bool r1 = lane & 1, r2 = lane & 2;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t t0 = o[0][i], t1 = o[1][i], t2 = o[2][i], t3 = o[3][i];
        const uint32_t u0 = r1 ? t1 : t0, u1 = r1 ? t2 : t1, u2 = r1 ? t3 : t2, u3 = r1 ? t0 : t3;
        o[0][i] = r2 ? u2 : u0; o[1][i] = r2 ? u3 : u1; o[2][i] = r2 ? u0 : u2; o[3][i] = r2 ? u1 : u3;
      }
      if (!(p.ablate & 1)) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          st_global_256(dst + 16 * ((j + lane) & 3), make_uint4(o[j][0], o[j][1], o[j][2], o[j][3]),
                        make_uint4(o[j][4], o[j][5], o[j][6], o[j][7]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == TC_EPI_WARPS + 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

inline bool configure_conv3x3_halo() {
  return cudaFuncSetAttribute(conv3x3_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, H3_SMEM) == cudaSuccess;
}

// true when the layer / call can run on the halo kernel
inline bool conv3x3_halo_ok(const TcWeights& t, int in_ld, int in_coff, int H, int W, int stride, const EpiParams& e) {
  return t.kind == TC_CONV && !t.transposed && t.kh == 3 && t.kw == 3 && t.cin == 128 && t.cout == 128 && t.bn == 128 &&
         t.n_tiles == 1 && t.b_box_rows == 128 && stride == 1 && H % 16 == 0 && W % 16 == 0 && in_ld % 8 == 0 &&
         in_coff % 8 == 0 && e.res1 == nullptr && e.res2 == nullptr && e.out_mode == OUT_T && e.out_c == 128 &&
         e.out_ld % 16 == 0 && e.out_coff % 16 == 0;
}

inline int launch_conv3x3_halo(TmaEncoder& tma, const TcWeights& t, const void* in, int in_ld, int in_coff, int H, int W,
                               int n, const EpiParams& e, int num_sms, int* errflag, cudaStream_t st, int* launches,
                               const Knobs& kn) {
  static thread_local std::map<TmapKey, CUtensorMap> cache;
  TmapKey key{in, in_ld, in_coff, t.cin, H, W, n, 1, H3_PITCH, H3_PITCH, 3000};
  auto it = cache.find(key);
  if (it == cache.end()) {
    CUtensorMap m;
    uint64_t dims[4] = {(uint64_t)t.cin, (uint64_t)W, (uint64_t)H, (uint64_t)n};
    uint64_t strides[3] = {(uint64_t)in_ld * 2, (uint64_t)W * in_ld * 2, (uint64_t)H * W * in_ld * 2};
    uint32_t box[4] = {TC_BK, H3_PITCH, H3_PITCH, 1};
    uint32_t es[4] = {1, 1, 1, 1};
    if (!tma.encode_h16(&m, (void*)((const h16*)in + in_coff), 4, dims, strides, box, es)) return -3;
    if (cache.size() > 1024) cache.clear();
    it = cache.emplace(key, m).first;
  }
  Halo3Params p;
  memset(&p, 0, sizeof p);
  p.n_img = n; p.H = H; p.W = W;
  p.tiles_x = W / 16; p.tiles_y = H / 16; p.total_tiles = n * p.tiles_x * p.tiles_y;
  p.bias = e.bias; p.act = e.act; p.out = e.out; p.out_ld = e.out_ld; p.out_coff = e.out_coff;
  p.ablate = kn.ablate; p.errflag = errflag;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = dim3((unsigned)(p.total_tiles < num_sms ? p.total_tiles : num_sms));
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = H3_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = kn.no_pdl ? 0 : 1;
  const cudaError_t le = cudaLaunchKernelEx(&cfg, conv3x3_halo_kernel, it->second, t.map, p);
  if (le != cudaSuccess) { tma.last_error = cudaGetErrorString(le); return -6; }
  (*launches)++;
  return 0;
}

}  // namespace bsr
