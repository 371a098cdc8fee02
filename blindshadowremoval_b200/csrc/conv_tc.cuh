// Implicit-GEMM convolution on tcgen05 tensor cores (h16 in, fp32 accumulate in TMEM) — persistent,
// warp-specialised, driven by a per-layer "step program".
//
// GEMM view:  D[128 pixels, N couts] = sum over steps  A_step[128, 64] . W_step[N, 64]^T
//   * A tile  = a BW x BH rectangle of GEMM-space pixels (BW*BH = 128) of one image, shifted by the
//     step's filter-tap offset (dy,dx), fetched by ONE 4-D TMA box {64 ch, BW, BH, 1} of the NHWC
//     input: out-of-bounds rows / columns / channels are zero-filled by the TMA unit, which is exactly
//     TF 'SAME' padding and the channel padding of the 99/257/261-wide tensors.  Stride-2 convs use
//     the same box with TMA element strides {1,2,2,1}.
//   * a step  = one A tile + the weight rows that multiply it + up to 2 MMAs into TMEM column ranges.
//       plain conv      : step = (tap, 64-channel block), one MMA of N = BN;
//       transposed conv : (model.py:153, 3x3 / stride 2 / SAME)  out[2i+py, 2j+px] = sum over taps
//                         kh = py, kw = px (mod 2) of in[i-(kh>>1), j-(kw>>1)] . W[kh,kw].  The 4
//                         sub-pixel phases share their A tiles, so a step = (shift, channel block) and
//                         its MMAs feed the accumulators [p10 | p00 | p01 | p11] of ALL phases:
//                         shift(0,0) -> 4 taps (N = 4*cout), (0,-1) -> 2, (-1,0) -> 2, (-1,-1) -> 1;
//       conv1 (7x7, 3ch): rows of the pre-packed image [N][H][W+8][8]: one step per filter ROW, whose
//                         64-element K block is the overlapping 8-pixel x 8-channel window;
//       heads (7x7, 64->2, model.py:204-205): "kw expansion": N = 7(kw) x 2 outputs, K = 7(kh) x 64, one
//                         image row (2 x 128 pixels) per tile; the horizontal taps are summed in the
//                         epilogue through shared memory, fused with the grey composition (model.py:246-252).
//   * W is packed once at load time (h16, K-major) and fetched by 2-D TMA; 128-byte swizzle on both
//     operands; UMMA M = 128, K = 16 per instruction.
// Persistent CTAs (one per SM) loop over tiles; accumulators are double-buffered in TMEM whenever
// 2 x columns <= 512, so the epilogue of tile t overlaps the TMA/MMA main loop of tile t+1.
// Warp roles (576 threads): warps 0-15 = epilogue (TMEM -> registers -> bias / residual adds / LeakyReLU ->
// global), warp 16 = TMA producer, warp 17 = TMEM allocator + MMA issuer.  The two single-thread roles get the
// HIGHEST warp ids on purpose: the warp arbiter favours high ids, and a starved MMA issuer stalls everything.
#pragma once
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "common.cuh"
#include "tc_common.cuh"

// Role timers (tools/role_timers.py) are compiled in only with -DBSR_ROLE_TIMERS: the clock reads sit on the
// single-thread producer / MMA critical paths.

namespace bsr {

constexpr int TC_BM = 128;          // pixels per A tile (UMMA M)
constexpr int TC_BK = 64;           // K elements per step (128 B of h16 = one swizzle row)
constexpr int TC_THREADS = 576;
constexpr int TC_EPI_WARPS = 16;
constexpr int TC_MAX_STEPS = 64;
constexpr int TC_SMEM_BUDGET = 200 * 1024;
// (the fp32 row-exchange buffers of round 1 are gone: horizontal taps are summed by warp shuffles)
constexpr int CLR_NB = 48;         // clr_conv1 accumulator group: 3 kw x 16 couts = 48 columns

enum TcKind : int { TC_CONV = 0, TC_CONVT_FUSED = 1, TC_ROWPACK = 2, TC_HEADS = 3, TC_CLR = 4, TC_PAIRX = 5 };
enum TcEpi : int { EPI_GENERIC = 0, EPI_HEADS = 1, EPI_CLR = 2, EPI_PLAIN = 3, EPI_QKV = 4 };

struct TcMma { int16_t col, n, brow; int16_t first; };   // first: bit 0 = overwrite (first tap), bit 1 = A sub-tile,
                                                         // bit 2 = A view starts one pixel row (128 B) into the tile (halo tiles)
struct TcStep {
  int8_t dy, dx;           // A tile shift in input pixels
  int8_t n_mma;
  int8_t n_a;              // bits 0-2: 1: one A tile; 2: two 64-wide K blocks (A sub-tiles at a_c0, a_c0+64, same MMAs over both);
                           // 3: two input ROWS (A sub-tiles at dy, dy+1), each MMA names its sub-tile (TcMma.first bit 1)
                           // bits 4-5: trailing 16-element K steps of the step's LAST K block that hold no input channel
                           // (cin = 257 pads to 320: the fifth block is one K step, not four) and are not issued
  int16_t a_c0;            // A channel coordinate (elements)
  int16_t b_rows;          // weight rows fetched for this step
  int32_t b_k;             // weight K coordinate (elements)
  int16_t b_row;           // first weight row (plus n-tile offset for plain convs)
  int16_t a_x0;            // extra GEMM-x offset of the A tile (heads: 0 / 128 for the two row halves)
  TcMma mma[8];
};

// Weights of the fused colour tail, passed BY VALUE as a kernel parameter: they live in the constant bank and are
// consumed directly as FFMA operands (no loads, no shared-memory traffic competing with the UMMA operand reads).
struct ClrWeights {
  float wg[9 * 16];      // gs taps of clr_conv1 [tap][out]
  float w2[16 * 16];     // clr_conv2 [in][out]
  float b2[16];
  float w3t[3 * 16];     // clr_conv3 transposed [out][in]
  float b3[4];
};

struct TcWeights {
  bool ready = false;
  int kind = TC_CONV;
  int kh = 0, kw = 0, cin = 0, cout = 0, transposed = 0;
  int cin_pad = 0;        // multiple of 64
  int bn = 0;             // accumulator columns owned by one n-tile (plain conv) / 4*cout (fused convT)
  int n_tiles = 0;
  int b_box_rows = 0;     // rows per weight TMA box
  int b_stage_rows = 0;   // max weight rows of one step (per K block)
  int a_sub = 1;          // max K blocks per step
  int n_steps = 0;
  int tile_w = 0;         // forced tile width (heads: 128 with two halves per tile), 0 = auto
  int rows_per_tile = 1;  // multi-row tiles: R output rows share their input rows (vertical-tap layers)
  int zero_acc = 0;       // see ConvTcParams::zero_acc
  int halves = 1;         // 2: a tile row is 256 pixels = two 128-pixel A tiles
  int b_resident = 0;     // whole weight matrix stays in shared memory for the life of the CTA
  int b_total_rows = 0;
  int b_res_kblocks = 1;
  int can_reside = 0;
  int can_pin = 0;        // n_tiles > 1: ONE n-tile's rows fit in shared memory -> a CTA may keep "its" n-tile resident
  TcStep steps[TC_MAX_STEPS];
  TcStep* steps_dev = nullptr;
  // halo variant of the fused transposed conv (tiles that are one image row of 128 pixels): the x-shifted A operands are
  // views of ONE 129-pixel tile, so a tile needs 2 TMA row loads instead of 4 (see pack_tc_weights)
  int n_steps_halo = 0;
  TcStep steps_halo[4];
  TcStep* steps_halo_dev = nullptr;
  h16* dev = nullptr;
  ClrWeights* clr = nullptr;   // EPI_CLR: host copy of the colour-tail weights (kernel parameter)
  CUtensorMap map;
  void release() {
    if (dev) cudaFree(dev);
    if (clr) delete clr;
    if (steps_dev) cudaFree(steps_dev);
    if (steps_halo_dev) cudaFree(steps_halo_dev);
    steps_halo_dev = nullptr;
    dev = nullptr;
    clr = nullptr;
    steps_dev = nullptr;
    ready = false;
  }
};

// Extra pointers of the fused epilogues.
struct EpiExtra {
  const float* img;        // [N,256,256,3] fp32 network input (grey reference)
  float* gs_out;           // optional user outputs
  float* mask22_out;
  float* difgs;            // fp32 [N,256,256] gs - grey (hole mask input)
  float* gs_f32;           // fp32 [N,256,256] gs (input of the colour tail)
  float* rgb_out;
  float* dif_out;
};

struct ConvTcParams {
  int n_img;
  int tiles_x, tiles_y;       // tiles per image in GEMM space
  int bw, bh;                 // A-tile rectangle
  int in_stride;              // 1 or 2: input coordinate = gemm coordinate * in_stride + shift
  int in_stride_x;            // x stride in units of the TMA x dimension (TC_PAIRX: pixel PAIRS -> 1)
  int out_scale;              // 1 (conv) or 2 (transposed)
  int OH, OW;
  int n_tiles;                // weight n-tiles (plain conv)
  int bn;                     // accumulator columns per tile
  int n_groups, group_cols;   // epilogue column groups (fused convT: 4 phases x cout)
  int group_phase[4];
  int n_steps;
  int b_box_rows;
  int stage_bytes, n_stages, acc_stages;
  int total_tiles;
  int epi_mode;
  int rows_per_tile, halves;  // multi-row tiles (groups = rows x halves)
  int pad_t, pad_l;           // SAME padding subtracted from the step shifts (plain convs)
  int a_sub;                  // A sub-tiles per stage (max n_a over the steps)
  int b_resident, b_total_rows;
  int b_res_rows;             // weight rows resident per CTA: all of them, or one n-tile's (b_pinned)
  int b_pinned;               // grid is a multiple of n_tiles, so CTA b only ever sees n-tile b % n_tiles and keeps just
                              // that tile's rows resident (qkv: 3 x 80 KB would not fit, 80 KB does)
  int b_res_kblocks;          // resident weights: number of 64-element K blocks kept ([kblock][row] layout)
  int st_chunk;               // staged TMA-store epilogue: 0 = direct per-thread stores, 16 = columns per warp and
                              // iteration (an iteration = 64 columns = two 32-channel SWIZZLE_64B boxes in smem)
  int st_bufs;                // staging buffers: 2 = double-buffered, 1 = single (one more barrier per iteration)
  int steps_bytes, epi_bytes; // shared-memory bytes of the step program / epilogue scratch (multiples of 128)
  int a_tile_bytes;           // bytes between the A sub-tiles of a stage: 16 KB, or 17 KB for 129-pixel halo tiles
  int a_box_bytes;            // bytes one A box delivers (expect_tx): 128 or 129 pixel rows of 128 B
  int b_kb_rows;              // streamed weights: rows between the K blocks of a step inside a stage
  int zero_acc;               // the epilogue leaves every accumulator stage zeroed, so NO MMA of the step program overwrites:
                              // the first tap of an output row rides in the same wide MMA as the rows already accumulating
  int ablate;                 // BSR_ABLATE (profiling only): 1 = no epilogue stores, 2 = no MMA, 4 = no A-tile TMA, 8 = timers
  long long* timers;          // [16] per launch (CTA 0): role wait / total cycle counters when ablate & 8
  int* errflag;
  const TcStep* steps;        // device copy of the step program
};

// vectorised helpers for the epilogue ------------------------------------------------------
__device__ __forceinline__ void add_res16(const void* base, size_t pix, int ld, int c, int climit, float* v) {
  const h16* p = (const h16*)base + pix * ld + c;
  if (c + 16 <= climit) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float2 t = unpack_h16x2(w[i]);
      v[2 * i] += t.x;
      v[2 * i + 1] += t.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (c + i < climit) v[i] += h16_to_f32(p[i]);
  }
}

// named barriers of the epilogue warps: 1 = all 512 epilogue threads, 2 + slot = the 256 threads of one row slot
__device__ __forceinline__ void epi_bar_all() { asm volatile("bar.sync 1, 512;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_slot(int slot) { asm volatile("bar.sync %0, 256;" ::"r"(2 + slot) : "memory"); }
// Programmatic dependent launch: everything before pdl_wait() (barrier init, TMEM alloc, weight / bias staging)
// overlaps the tail of the previous kernel in the stream; nothing produced by that kernel is touched before it.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ void add_h16x16(const uint4& a, const uint4& b, float* v) {
  const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float2 t = unpack_h16x2(w[i]);
    v[2 * i] += t.x;
    v[2 * i + 1] += t.y;
  }
}

template <int EPI, bool RES>
__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                const __grid_constant__ CUtensorMap tmB,
                                                                const __grid_constant__ ConvTcParams p,
                                                                const EpiParams e, const EpiExtra x,
                                                                const __grid_constant__ ClrWeights cw,
                                                                const __grid_constant__ CUtensorMap tmO) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_bytes = (uint32_t)p.a_tile_bytes;
  const uint32_t stage_bytes = (uint32_t)p.stage_bytes;
  const uint32_t sBres = smem_base + (uint32_t)p.n_stages * stage_bytes;            // resident weights (optional)
  const uint32_t bars = sBres + (p.b_resident ? (uint32_t)(p.b_res_kblocks * p.b_res_rows) * 128u : 0u);
  // full[8], empty[8], tmem_full[2], tmem_empty[2], bres, tmem slot
  const uint32_t bar_full = bars, bar_empty = bars + 64, bar_tfull = bars + 128, bar_tempty = bars + 144;
  const uint32_t bar_bres = bars + 160, tmem_slot = bars + 168, bar_zero = bars + 176;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_al + (tmem_slot - smem_base));
  TcStep* steps = reinterpret_cast<TcStep*>(smem_al + (bars + 192 - smem_base));
  float* epi_smem = reinterpret_cast<float*>(smem_al + (bars + 192 + (uint32_t)p.steps_bytes - smem_base));
  // output staging of the TMA-store epilogue: st_bufs buffers of 128 rows x 64 h16 columns
  const uint32_t st_stage = (bars + 192 + (uint32_t)p.steps_bytes + (uint32_t)p.epi_bytes + 511u) & ~511u;   // 512 B = one 64B-swizzle atom
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(p.steps);
    uint32_t* dst = reinterpret_cast<uint32_t*>(steps);
    for (int i = threadIdx.x; i < p.n_steps * (int)(sizeof(TcStep) / 4); i += blockDim.x) dst[i] = src[i];
  }

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int n_tiles = p.n_tiles, total_tiles = p.total_tiles, n_steps = p.n_steps;
  const uint32_t n_stages = (uint32_t)p.n_stages, acc_stages = (uint32_t)p.acc_stages;
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)p.bn * acc_stages) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    if (p.st_chunk) prefetch_tmap(&tmO);
    for (int s = 0; s < p.n_stages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_tfull + 8 * s, 1);
      mbar_init(bar_tempty + 8 * s, TC_EPI_WARPS);
    }
    mbar_init(bar_bres, 1);
    mbar_init(bar_zero, TC_EPI_WARPS);
    fence_barrier_init();
  }
  if (warp == TC_EPI_WARPS + 1) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_launch_dependents();

  if (warp == TC_EPI_WARPS) {
    // ================= TMA producer (whole warp converged; one elected lane issues) =================
    {
      uint32_t it = 0, s = 0, ph = 0;          // stage / phase kept incrementally (no runtime division on the issue path)
      bool ok = true;
      const bool leader = elect_one();
      const int in_stride = p.in_stride, in_stride_x = p.in_stride_x, pad_l = p.pad_l, pad_t = p.pad_t, b_resident = p.b_resident;
      const int b_box_rows = p.b_box_rows, bn = p.bn, ablate = p.ablate;
      const uint32_t a_sub = (uint32_t)p.a_sub, b_kb_stride = (uint32_t)p.b_kb_rows;
      const int tiles_x = p.tiles_x, tile_h = p.bh * p.rows_per_tile, tile_w = p.bw * p.halves;
      if (b_resident && leader) {          // weights are static: fetched before the dependency wait
        const int row0 = p.b_pinned ? (int)(blockIdx.x % (unsigned)p.n_tiles) * bn : 0;
        mbar_expect_tx(bar_bres, (uint32_t)(p.b_res_kblocks * p.b_res_rows) * 128u);
        for (int kb = 0; kb < p.b_res_kblocks; ++kb)
          for (int r = 0; r < p.b_res_rows; r += b_box_rows)
            tma_load_2d(sBres + (uint32_t)(kb * p.b_res_rows + r) * 128u, &tmB, bar_bres, kb * TC_BK, row0 + r);
      }
      const bool tm = (ablate & 8) && blockIdx.x == 0 && leader;
      long long t_wait = 0, t_dep = 0, t_tma = 0;
      const long long t_start = BSR_CLK();
      pdl_wait();
      t_dep = BSR_CLK() - t_start;
      for (int tile = blockIdx.x; tile < total_tiles && ok; tile += gridDim.x) {
        const int ntile = tile % n_tiles, mt = tile / n_tiles;
        const int n = mt / tiles_per_img, tr = mt % tiles_per_img;
        const int gy0 = (tr / tiles_x) * tile_h, gx0 = (tr % tiles_x) * tile_w;
        const int xbase = gx0 * in_stride_x - pad_l, ybase = gy0 * in_stride - pad_t, brow_base = ntile * bn;
        for (int si = 0; si < n_steps; ++si, ++it) {
          const TcStep& sp = steps[si];
          const long long tw = BSR_CLK();
          ok = mbar_wait(bar_empty + 8 * s, ph ^ 1u, p.errflag, 1);
          const long long tp0 = BSR_CLK();
          t_wait += tp0 - tw;
          if (!ok) break;
          const uint32_t dstA = smem_base + s * stage_bytes, dstB = dstA + a_sub * a_bytes;
          if (leader) {
            // n_a: 1 one A tile; 2 two K blocks (channels + 64); 3 two input rows; 4 two K blocks from x-neighbour tiles;
            // 5 the same from ONE 129-row halo tile (the second block is a view one row in)
            const uint32_t amode = (uint32_t)sp.n_a & 7u;
            const uint32_t rowmode = amode == 3, xmode = amode == 4, hmode = amode == 5,
                           na = rowmode ? 1u : ((xmode || hmode) ? 2u : amode), nsub = (amode >= 2 && !hmode) ? 2u : 1u;
            const uint32_t bbytes = b_resident ? 0u : na * (uint32_t)sp.b_rows * 128u;
            if (ablate & 4) {
              mbar_expect_tx(bar_full + 8 * s, bbytes);
            } else {
              mbar_expect_tx(bar_full + 8 * s, nsub * (uint32_t)p.a_box_bytes + bbytes);
              const int ax = xbase + sp.a_x0 * in_stride_x + sp.dx, ay = ybase + sp.dy;
              tma_load_4d(dstA, &tmA, bar_full + 8 * s, sp.a_c0, ax, ay, n);
              if (nsub == 2)
                tma_load_4d(dstA + a_bytes, &tmA, bar_full + 8 * s, sp.a_c0 + ((rowmode || xmode) ? 0 : TC_BK), ax + (int)xmode,
                            ay + (int)rowmode, n);
            }
            if (!b_resident) {
              const int row0 = sp.b_row + brow_base;
              for (uint32_t kb = 0; kb < na; ++kb)
                for (int r = 0; r < sp.b_rows; r += b_box_rows)
                  tma_load_2d(dstB + (kb * b_kb_stride + (uint32_t)r) * 128u, &tmB, bar_full + 8 * s,
                              sp.b_k + (int)kb * TC_BK, row0 + r);
            }
          }
          __syncwarp();
          t_tma += BSR_CLK() - tp0;
          if (++s == n_stages) { s = 0; ph ^= 1u; }
        }
      }
      if (tm) { p.timers[15] = t_tma; p.timers[0] = BSR_CLK() - t_start; p.timers[1] = t_wait; p.timers[2] = t_dep; p.timers[3] = it; }
    }
  } else if (warp == TC_EPI_WARPS + 1) {
    // ================= MMA issuer (whole warp converged; one elected lane issues) =================
    {
      uint32_t it = 0, tcount = 0, s = 0, ph = 0, as = 0, aph = 0;
      bool ok = true;
      const bool leader = elect_one();
      const int b_resident = p.b_resident, b_total_rows = p.b_res_rows, bn = p.bn, ablate = p.ablate;
      const uint32_t a_sub = (uint32_t)p.a_sub;
      // distance (in 16-byte units) between the weight rows of the two K blocks of a step
      const uint32_t b_kb_lo = b_resident ? (uint32_t)b_total_rows * 8u : (uint32_t)p.b_kb_rows * 8u;
      const bool tm = (ablate & 8) && blockIdx.x == 0 && leader;
      long long t_wfull = 0, t_wtempty = 0, t_fence = 0, t_issue = 0, t_commit = 0, t_tcommit = 0, t_gap_s = 0, t_gap_t = 0, t_prev = 0;
      const long long t_start = BSR_CLK();
      if (b_resident) ok = mbar_wait(bar_bres, 0, p.errflag, 5);
      if (ok && p.zero_acc) ok = mbar_wait(bar_zero, 0, p.errflag, 6);      // both accumulator stages start out zeroed
      const long long t_res = BSR_CLK() - t_start;
      const uint32_t idesc_m = umma_idesc_h16(TC_BM, 0);
      for (int tile = blockIdx.x; tile < total_tiles && ok; tile += gridDim.x, ++tcount) {
        const long long tw0 = BSR_CLK();
        if (tcount) t_gap_t += tw0 - t_prev;
        ok = mbar_wait(bar_tempty + 8 * as, aph ^ 1u, p.errflag, 4, true);
        const long long tw0b = BSR_CLK();
        t_wtempty += tw0b - tw0;
        if (!ok) break;
        tc_fence_after();
        t_prev = BSR_CLK();
        t_fence += t_prev - tw0b;
        const uint32_t acc = tmem_base + as * (uint32_t)bn;
        const int brow_base = p.b_pinned ? 0 : (tile % n_tiles) * bn;
        for (int si = 0; si < n_steps; ++si, ++it) {
          const TcStep& sp = steps[si];
          const long long tw1 = BSR_CLK();
          t_gap_s += tw1 - t_prev;
          ok = mbar_wait(bar_full + 8 * s, ph, p.errflag, 2, true);
          const long long tc0 = BSR_CLK();
          t_wfull += tc0 - tw1;
          if (!ok) break;
          tc_fence_after();
          const long long tc1 = BSR_CLK();
          t_fence += tc1 - tc0;
          const uint32_t sA = smem_base + s * stage_bytes;
          const uint32_t sB = b_resident ? sBres + (uint32_t)((sp.b_k >> 6) * b_total_rows + sp.b_row + brow_base) * 128u
                                         : sA + a_sub * a_bytes;
          const uint32_t a_lo = umma_desc_lo(sA), b_lo0 = umma_desc_lo(sB);
          const int nm = (ablate & 2) ? 0 : sp.n_mma;
          const uint32_t amode = (uint32_t)sp.n_a & 7u;
          const uint32_t na = amode >= 4 ? 2u : amode;          // 4 / 5 = two K blocks whose A tiles are x-neighbours
          // SKIP = false is the common program (all four K steps of every block); SKIP = true drops the all-padding K steps
          // at the end of the step's last K block.  Two instantiations, selected once per step: the issue loop of the
          // small-N layers is on the critical path and does not tolerate per-MMA conditionals (clr_conv1 +15 % with them).
          auto issue = [&](auto skip_tag) {
            constexpr bool SKIP = decltype(skip_tag)::value;
            const uint32_t nk_last = SKIP ? 4u - (((uint32_t)sp.n_a >> 4) & 3u) : 4u, nk_first = na == 2 ? 4u : nk_last;
            for (int m = 0; m < nm; ++m) {
              const TcMma mm = sp.mma[m];
              const uint32_t b_lo = b_lo0 + (uint32_t)mm.brow * 8u;          // 128-byte rows, address >> 4
              const uint32_t idesc = idesc_m | ((uint32_t)(mm.n >> 3) << 17);
              const uint32_t d = acc + (uint32_t)mm.col;
              // halo view: the operand starts one 128-byte pixel row into the tile.  The 128B swizzle is a function of the
              // shared-memory ADDRESS bits (measured: matrix base offset 0 is right, 1 is wrong - tools/halo_probe.py), so a
              // view that starts anywhere on a 128-byte row boundary of a TMA-written tile reads the right elements.
              const uint32_t voff = (mm.first & 4) ? 8u : 0u;
              const uint32_t a0 = a_lo + ((mm.first & 2) ? (a_bytes >> 4) : 0u) + voff;     // row mode: second input row
              umma_h16_lo(d, a0, b_lo, idesc, (mm.first & 1) ? 0u : 1u);
              if (!SKIP || nk_first > 1) umma_h16_lo(d, a0 + 2, b_lo + 2, idesc, 1u);
              if (!SKIP || nk_first > 2) umma_h16_lo(d, a0 + 4, b_lo + 4, idesc, 1u);
              if (!SKIP || nk_first > 3) umma_h16_lo(d, a0 + 6, b_lo + 6, idesc, 1u);
              if (na == 2) {       // second 64-wide K block of the step
                // n_a = 5: the second K block is the x-neighbour pixel pair = the same halo tile one row further
                const uint32_t a1 = a_lo + (amode == 5 ? 8u : (a_bytes >> 4)) + voff, b1 = b_lo + b_kb_lo;
                umma_h16_lo(d, a1, b1, idesc, 1u);
                if (!SKIP || nk_last > 1) umma_h16_lo(d, a1 + 2, b1 + 2, idesc, 1u);
                if (!SKIP || nk_last > 2) umma_h16_lo(d, a1 + 4, b1 + 4, idesc, 1u);
                if (!SKIP || nk_last > 3) umma_h16_lo(d, a1 + 6, b1 + 6, idesc, 1u);
              }
            }
          };
          if (leader) {
            if ((uint32_t)sp.n_a >> 4) issue(std::true_type{});
            else issue(std::false_type{});
          }
          const long long tc2 = BSR_CLK();
          t_issue += tc2 - tc1;
          if (leader) umma_commit(bar_empty + 8 * s);
          __syncwarp();
          t_prev = BSR_CLK();
          t_commit += t_prev - tc2;
          if (++s == n_stages) { s = 0; ph ^= 1u; }
        }
        const long long tt0 = BSR_CLK();
        if (leader) umma_commit(bar_tfull + 8 * as);
        __syncwarp();
        t_prev = BSR_CLK();
        t_tcommit += t_prev - tt0;
        if (++as == acc_stages) { as = 0; aph ^= 1u; }
      }
      if (tm) { p.timers[12] = (ablate & 16) ? t_gap_s : t_fence; p.timers[13] = t_issue; p.timers[14] = (ablate & 16) ? t_gap_t : t_commit; }
      if (tm) { p.timers[4] = BSR_CLK() - t_start; p.timers[5] = t_wfull; p.timers[6] = t_wtempty; p.timers[7] = t_res + t_tcommit; p.timers[8] = tcount; }
    }
  } else {
    // ================= epilogue: 16 warps.  warp ew reads TMEM lanes 32*(warp%4).. (hardware rule) =================
    const int ew = warp;
    const int q = warp & 3;                       // TMEM lane quarter = rows 32q .. 32q+31 of the A tile
    const int cg = ew >> 2;                       // 0..3: column-chunk group (GENERIC) / row slot + half (HEADS, CLR)
    const int r = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const int ethread = threadIdx.x;
    uint32_t tcount = 0, st_it = 0;
    bool ok = true;
    // static data first (bias, colour-tail weights): not produced by the previous kernel
    float* bias_s = epi_smem;                   // [512] (the device bias buffer is zero-padded past cout)
    float* epi_work = epi_smem + 512;
    for (int i = ethread; i < 512 && i * 4 < p.epi_bytes; i += TC_EPI_WARPS * 32) bias_s[i] = __ldg(e.bias + i);
    if (p.zero_acc) {
      // zero this warp's share of the allocated TMEM columns (4 warps per lane quarter: column chunks cg, cg + 4, ...)
      for (uint32_t c = (uint32_t)cg * 16u; c < tmem_cols; c += 64u) tmem_zero16(tmem_base + lane_addr + c);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_zero);
    }
    epi_bar_all();
    pdl_wait();
    const int OH = p.OH, OW = p.OW, bn = p.bn, rows_per_tile = p.rows_per_tile;
    const bool tm = (p.ablate & 8) && blockIdx.x == 0 && ethread == 0;
    long long t_wtfull = 0;
    const long long t_start = BSR_CLK();
    for (int tile = blockIdx.x; tile < total_tiles && ok; tile += gridDim.x, ++tcount) {
      const int ntile = tile % n_tiles, mt = tile / n_tiles;
      const int n = mt / tiles_per_img, tr = mt % tiles_per_img;
      const uint32_t as = acc_stages == 2 ? (tcount & 1u) : 0u, aph = (acc_stages == 2 ? (tcount >> 1) : tcount) & 1u;
      const uint32_t acc = tmem_base + as * (uint32_t)bn + lane_addr;
      if (EPI == EPI_GENERIC && !RES && p.st_chunk) {
        // ---- staged epilogue: accumulators -> h16 -> 64B-swizzled shared memory -> TMA store.  Per-thread-row global
        // stores touch 32 cache lines per warp instruction and (measured) slow the UMMA operand
        // reads that share the L1/shared-memory pipe; the bulk store reads the staged tile at full line width instead.
        const int group_cols = p.group_cols, total_cols = p.n_groups * group_cols;
        constexpr int cw_ = 16, iter_cols = 4 * cw_;
        const uint32_t buf_bytes = (uint32_t)iter_cols * 256u;                 // 128 rows x iter_cols h16
        const int gx0 = (tr % p.tiles_x) * p.bw, gy0 = (tr / p.tiles_x) * p.bh * rows_per_tile;
        ok = mbar_wait(bar_tfull + 8 * as, aph, p.errflag, 3);
        if (!ok) break;
        tc_fence_after();
        const bool single = p.st_bufs == 1;
        for (int col0 = 0; col0 < total_cols; col0 += iter_cols, ++st_it) {
          const uint32_t buf = st_stage + (single ? 0u : (st_it & 1u) * buf_bytes);
          const int colw = cg * cw_, col = col0 + colw;                        // this warp's columns of the iteration
          uint32_t o[8];
          if (col < total_cols) {
            float v[16];
            tmem_ld16(acc + (uint32_t)col, v);
            if (p.zero_acc) tmem_zero16(acc + (uint32_t)col);
            const int j = col % group_cols;                                    // channel inside the group
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              const float4 b4 = *reinterpret_cast<const float4*>(bias_s + j + i);
              v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
            }
            if (e.act) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], kLeaky * v[i]);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = pack_h16x2(v[2 * i], v[2 * i + 1]);
          }
          if (col0 + iter_cols >= total_cols) {          // accumulators drained: the MMA warp may start the next tile
            if (p.zero_acc) tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8 * as);
          }
          if (single) {       // one buffer: the previous bulk store must have finished reading it before anyone writes
            if (ethread == 0) bulk_wait_read0();
            epi_bar_all();
          }
          if (col < total_cols) {
            // box = 32 channels (64-byte rows); 16-byte piece `pc` of row r lands at r*64 + ((pc ^ (r>>1)) & 3)*16
            const uint32_t box = buf + (uint32_t)(colw >> 5) * 8192u;
            const uint32_t pc0 = (uint32_t)(colw & 31) >> 3;
            const uint32_t rowb = box + (uint32_t)r * 64u, sw = ((uint32_t)r >> 1) & 3u;
            {
              const uint32_t a = rowb + ((pc0 ^ sw) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
            }
            {
              const uint32_t a = rowb + (((pc0 + 1u) ^ sw) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]) : "memory");
            }
          }
          fence_async_smem();
          // double-buffered: the bulk store issued one iteration ago has finished READING the other buffer before anyone
          // passes the barrier below and starts overwriting it
          if (!single && ethread == 0) bulk_wait_read0();
          epi_bar_all();
          if (ethread == 0 && !(p.ablate & 1)) {
            for (int b = 0; b * 32 < iter_cols; ++b) {
              const int bc = col0 + b * 32;
              if (bc < total_cols) {
                const int g = bc / group_cols, j = bc - g * group_cols;
                const int phase = rows_per_tile > 1 ? 0 : p.group_phase[g];
                const int gy = gy0 + (rows_per_tile > 1 ? g * p.bh : 0);
                tma_store_5d(&tmO, buf + (uint32_t)b * 8192u, e.out_coff + j, phase & 1, gx0, phase >> 1,
                             n * (OH / p.out_scale) + gy);
              }
            }
            bulk_commit();
          }
        }
      } else if (EPI == EPI_GENERIC) {
        const int bw = p.bw, bh = p.bh, out_scale = p.out_scale, group_cols = p.group_cols;
        const int gyb = (tr / p.tiles_x) * bh * rows_per_tile + r / bw, gx = (tr % p.tiles_x) * bw + r % bw;
        const bool vec_ok = (e.out_ld % 8 == 0) && (e.out_coff % 8 == 0);
        // 32-byte aligned 16-channel chunks (all chunk starts are multiples of 16 channels): one 256-bit store per chunk
        const bool vec32_ok = e.out_mode == OUT_QKV || ((e.out_ld % 16 == 0) && (e.out_coff % 16 == 0));
        const int cbase = p.n_groups == 1 ? ntile * bn : 0;
        const int total_cols = p.n_groups * group_cols;
        // residual / skip operands of this thread's first 3 chunks are fetched BEFORE the accumulator wait, so their
        // global-load latency hides behind the MMA main loop (plain convs only: one pixel per thread)
        uint4 rb1[3][2], rb2[3][2];
        uint32_t m1 = 0, m2 = 0;
        if (RES) {
          const size_t pix0 = ((size_t)n * OH + gyb) * OW + gx;
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const int col = (4 * k + cg) * 16, c = cbase + col;
            if (col < bn && c + 16 <= e.out_c) {
              if (e.res1 != nullptr && c + 16 <= e.res1_c) {
                const h16* src = (const h16*)e.res1 + pix0 * e.res1_ld + c;
                if (e.res1_ld % 16 == 0) ld_global_256(src, rb1[k][0], rb1[k][1]);
                else { rb1[k][0] = reinterpret_cast<const uint4*>(src)[0]; rb1[k][1] = reinterpret_cast<const uint4*>(src)[1]; }
                m1 |= 1u << k;
              }
              if (e.res2 != nullptr && c + 16 <= e.res2_c) {
                const h16* src = (const h16*)e.res2 + pix0 * e.res2_ld + c;
                if (e.res2_ld % 16 == 0) ld_global_256(src, rb2[k][0], rb2[k][1]);
                else { rb2[k][0] = reinterpret_cast<const uint4*>(src)[0]; rb2[k][1] = reinterpret_cast<const uint4*>(src)[1]; }
                m2 |= 1u << k;
              }
            }
          }
        }
        const long long tw2 = BSR_CLK();
        ok = mbar_wait(bar_tfull + 8 * as, aph, p.errflag, 3);
        t_wtfull += BSR_CLK() - tw2;
        if (!ok) break;
        tc_fence_after();
        // running (group, column-in-group) of this warp's chunk; chunks advance by 4 x 16 columns
        int g = 0, j = cg * 16;
        while (j >= group_cols) { j -= group_cols; ++g; }
        int g_cur = -1;
        size_t pix = 0;
        int gy = 0;
        auto chunk = [&](const int col, const bool pf1, const uint4& r1a, const uint4& r1b, const bool pf2,
                         const uint4& r2a, const uint4& r2b) {
          if (g != g_cur) {          // per-group output pixel (sub-pixel phase or row), computed once per group
            g_cur = g;
            const int phase = rows_per_tile > 1 ? 0 : p.group_phase[g];
            gy = gyb + (rows_per_tile > 1 ? g * bh : 0);
            const int oy = gy * out_scale + (phase >> 1), ox = gx * out_scale + (phase & 1);
            pix = ((size_t)n * OH + oy) * OW + ox;
          }
          const int c = cbase + j;
          if (p.zero_acc && c >= e.out_c) tmem_zero16(acc + (uint32_t)col);
          if (c < e.out_c) {
            float v[16];
            tmem_ld16(acc + (uint32_t)col, v);
            if (p.zero_acc) tmem_zero16(acc + (uint32_t)col);
            if (c + 16 <= e.out_c && (vec_ok || e.out_mode != OUT_T)) {
#pragma unroll
              for (int i = 0; i < 16; i += 4) {
                const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c + i);
                v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
              }
              if (RES) {
                if (pf1) add_h16x16(r1a, r1b, v);
                else if (e.res1 != nullptr && c < e.res1_c) add_res16(e.res1, pix, e.res1_ld, c, e.res1_c, v);
                if (pf2) add_h16x16(r2a, r2b, v);
                else if (e.res2 != nullptr && c < e.res2_c) add_res16(e.res2, pix, e.res2_ld, c, e.res2_c, v);
              }
              if (e.act) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], kLeaky * v[i]);      // leaky, alpha < 1
              }
              if (e.out_mode == OUT_F32) {
                float* dst = (float*)e.out + pix * e.out_ld + e.out_coff + c;
#pragma unroll
                for (int i = 0; i < 16; ++i) dst[i] = v[i];
              } else if (e.out_mode == OUT_QKV && c >= 256 && !e.v_natural) {
                // g -> V^T[n][c-256][s]: for a fixed channel the 32 lanes write consecutive tokens
                h16* dst = (h16*)e.out2 + ((size_t)n * 128 + (c - 256)) * e.spatial + (gy * OW + gx);
#pragma unroll
                for (int i = 0; i < 16; ++i) dst[(size_t)i * e.spatial] = f32_to_h16(v[i]);
              } else {
                uint4 o0, o1;
                o0.x = pack_h16x2(v[0], v[1]); o0.y = pack_h16x2(v[2], v[3]);
                o0.z = pack_h16x2(v[4], v[5]); o0.w = pack_h16x2(v[6], v[7]);
                o1.x = pack_h16x2(v[8], v[9]); o1.y = pack_h16x2(v[10], v[11]);
                o1.z = pack_h16x2(v[12], v[13]); o1.w = pack_h16x2(v[14], v[15]);
                const size_t off = e.out_mode == OUT_QKV ? pix * 256 + c : pix * e.out_ld + e.out_coff + c;
                // OUT_QKV with natural V: channels >= 256 go to V[pix][c - 256] (same vector store, other base)
                uint4* dst = (e.out_mode == OUT_QKV && c >= 256)
                                 ? reinterpret_cast<uint4*>((h16*)e.out2 + pix * 128 + (c - 256))
                                 : reinterpret_cast<uint4*>((h16*)e.out + off);
                if (!(p.ablate & 1)) {
                  if (vec32_ok) {
                    st_global_256(dst, o0, o1);
                  } else {
                    dst[0] = o0;
                    dst[1] = o1;
                  }
                }
              }
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) epi_store<h16>(e, pix, c + i, v[i]);
            }
          }
          j += 64;
          while (j >= group_cols) { j -= group_cols; ++g; }
        };
        int col = cg * 16;
        if (!RES && p.n_groups == 4 && e.out_mode == OUT_T && vec32_ok && group_cols == 64 && e.out_c == 64 && bn == 256) {
          // Fused 4-phase transposed conv with 64 output channels whose stores cannot be staged (up3 / clr_up3: 144 KB of
          // resident weights): this warp owns ALL 64 channels of ONE sub-pixel phase (accumulator columns [64 cg, 64 cg + 64)),
          // i.e. one whole 128-byte output pixel per lane.  The four 16-column chunks are fetched with ONE TMEM round trip
          // and the accumulators are released before the stores.  Every chunk leaves as one 256-bit store, and lane L stores
          // chunk (j + L) % 4 in store j: a warp-wide store whose 32 lanes all write the SAME 32-byte slice of their
          // (different) 128-byte lines costs 60 L1 data-pipe wavefronts, with the four slices spread over the lanes it costs
          // 16 (tools/micro/store_pattern.cu, profiles/r2_store_pattern.txt) - and that pipe also feeds the MMA operands.
          float v[4][16];
#pragma unroll
          for (int k = 0; k < 4; ++k) tmem_ld16_nowait(acc + (uint32_t)(cg * 64 + 16 * k), v[k]);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_tempty + 8 * as);
          const int phase = p.group_phase[cg];
          const int oy = gyb * out_scale + (phase >> 1), ox = gx * out_scale + (phase & 1);
          h16* const dst = (h16*)e.out + (((size_t)n * OH + oy) * OW + ox) * e.out_ld + e.out_coff;
          uint32_t o[4][8];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              const float4 b4 = *reinterpret_cast<const float4*>(bias_s + 16 * k + i);
              float a0 = v[k][i] + b4.x, a1 = v[k][i + 1] + b4.y, a2 = v[k][i + 2] + b4.z, a3 = v[k][i + 3] + b4.w;
              if (e.act) {
                a0 = fmaxf(a0, kLeaky * a0); a1 = fmaxf(a1, kLeaky * a1);
                a2 = fmaxf(a2, kLeaky * a2); a3 = fmaxf(a3, kLeaky * a3);
              }
              o[k][i >> 1] = pack_h16x2(a0, a1);
              o[k][(i >> 1) + 1] = pack_h16x2(a2, a3);
            }
          }
          // rotate the four chunks by (lane % 4) in two conditional stages: afterwards o[j] holds chunk (j + lane) % 4
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
          continue;
        }
        if (!RES && total_cols == 128 && p.n_groups == 1 && rows_per_tile == 1 && !p.zero_acc && vec32_ok &&
            e.out_mode != OUT_F32 && cbase + 128 <= e.out_c && (e.out_mode != OUT_QKV || cbase < 256 || e.v_natural)) {
          // 128-column tiles with direct stores (the q / k / v conv): this warp owns the 32 columns [32 cg, 32 cg + 32) and lane
          // L stores chunk (j + L) % 2 in store j, so the lanes of one warp-wide store write two different 32-byte slices of
          // their lines instead of one (see the fused transposed-conv path above for the L1 data-pipe cost of the pattern).
          float v[2][16];
          tmem_ld16_nowait(acc + (uint32_t)(cg * 32), v[0]);
          tmem_ld16_nowait(acc + (uint32_t)(cg * 32 + 16), v[1]);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_tempty + 8 * as);
          const int c0 = cbase + cg * 32;
          const int phase = p.group_phase[0];          // one launch per sub-pixel phase for wide transposed convs
          const size_t pix1 = ((size_t)n * OH + gyb * out_scale + (phase >> 1)) * OW + gx * out_scale + (phase & 1);
          h16* const dst = e.out_mode == OUT_QKV ? (c0 >= 256 ? (h16*)e.out2 + pix1 * 128 + (c0 - 256) : (h16*)e.out + pix1 * 256 + c0)
                                                 : (h16*)e.out + pix1 * e.out_ld + e.out_coff + c0;
          uint32_t o[2][8];
#pragma unroll
          for (int k = 0; k < 2; ++k) {
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c0 + 16 * k + i);
              float a0 = v[k][i] + b4.x, a1 = v[k][i + 1] + b4.y, a2 = v[k][i + 2] + b4.z, a3 = v[k][i + 3] + b4.w;
              if (e.act) {
                a0 = fmaxf(a0, kLeaky * a0); a1 = fmaxf(a1, kLeaky * a1);
                a2 = fmaxf(a2, kLeaky * a2); a3 = fmaxf(a3, kLeaky * a3);
              }
              o[k][i >> 1] = pack_h16x2(a0, a1);
              o[k][(i >> 1) + 1] = pack_h16x2(a2, a3);
            }
          }
          const bool r1 = lane & 1;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint32_t t0 = o[0][i], t1 = o[1][i];
            o[0][i] = r1 ? t1 : t0; o[1][i] = r1 ? t0 : t1;
          }
          if (!(p.ablate & 1)) {
#pragma unroll
            for (int j = 0; j < 2; ++j)
              st_global_256(dst + 16 * ((j + lane) & 1), make_uint4(o[j][0], o[j][1], o[j][2], o[j][3]),
                            make_uint4(o[j][4], o[j][5], o[j][6], o[j][7]));
          }
          continue;
        }
        if (RES) {
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            if (col < total_cols) chunk(col, (m1 >> k) & 1u, rb1[k][0], rb1[k][1], (m2 >> k) & 1u, rb2[k][0], rb2[k][1]);
            col += 64;
          }
        }
#pragma unroll 1
        for (; col < total_cols; col += 64) chunk(col, false, rb1[0][0], rb1[0][1], false, rb2[0][0], rb2[0][1]);
        if (p.zero_acc) tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tempty + 8 * as);
      } else if (EPI == EPI_QKV) {
        // q / k / v projection (OUT_QKV, natural V) in two 192-column tiles; its own instantiation like EPI_PLAIN
        const int bw = p.bw, bh = p.bh;
        const int gyb = (tr / p.tiles_x) * bh * rows_per_tile + r / bw, gx = (tr % p.tiles_x) * bw + r % bw;
        const int cbase = ntile * bn;
        ok = mbar_wait(bar_tfull + 8 * as, aph, p.errflag, 3);
        if (!ok) break;
        tc_fence_after();
        {
          // q / k / v projection in two 192-column tiles: this warp owns 48 columns = three 16-channel chunks, fetched with
          // one TMEM round trip.  Lane L stores chunk (j + L) % 3 in store j: theta|phi pixels are 512 bytes and g pixels
          // 256 bytes apart, so without the rotation all 32 lanes of a store would hit the same 32-byte slice of their lines.
          const int c0 = cbase + 48 * cg;
          float v[3][16];
#pragma unroll
          for (int k = 0; k < 3; ++k) tmem_ld16_nowait(acc + (uint32_t)(48 * cg + 16 * k), v[k]);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_tempty + 8 * as);
          uint32_t o[3][8];
#pragma unroll
          for (int k = 0; k < 3; ++k) {
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c0 + 16 * k + i);
              float a0 = v[k][i] + b4.x, a1 = v[k][i + 1] + b4.y, a2 = v[k][i + 2] + b4.z, a3 = v[k][i + 3] + b4.w;
              if (e.act) {
                a0 = fmaxf(a0, kLeaky * a0); a1 = fmaxf(a1, kLeaky * a1);
                a2 = fmaxf(a2, kLeaky * a2); a3 = fmaxf(a3, kLeaky * a3);
              }
              o[k][i >> 1] = pack_h16x2(a0, a1);
              o[k][(i >> 1) + 1] = pack_h16x2(a2, a3);
            }
          }
          const int r3 = lane % 3;
          const bool z0 = r3 == 0, z1 = r3 == 1;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint32_t t0 = o[0][i], t1 = o[1][i], t2 = o[2][i];
            o[0][i] = z0 ? t0 : (z1 ? t1 : t2);
            o[1][i] = z0 ? t1 : (z1 ? t2 : t0);
            o[2][i] = z0 ? t2 : (z1 ? t0 : t1);
          }
          const size_t pixq = ((size_t)n * OH + gyb) * OW + gx;
          if (!(p.ablate & 1)) {
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              int m = j + r3;
              if (m >= 3) m -= 3;
              const int c = c0 + 16 * m;
              h16* const dq = c >= 256 ? (h16*)e.out2 + pixq * 128 + (c - 256) : (h16*)e.out + pixq * 256 + c;
              st_global_256(dq, make_uint4(o[j][0], o[j][1], o[j][2], o[j][3]), make_uint4(o[j][4], o[j][5], o[j][6], o[j][7]));
            }
          }
        }
      } else if (EPI == EPI_PLAIN) {
        // Lean direct-store epilogue for plain convs (selected by launch_conv_tc: no residual, one column group, 16-channel
        // aligned h16 output).  Its own kernel instantiation, so its registers do not compete with the general paths.
        const int bw = p.bw, bh = p.bh, out_scale = p.out_scale;
        const int gyb = (tr / p.tiles_x) * bh * rows_per_tile + r / bw, gx = (tr % p.tiles_x) * bw + r % bw;
        const int cbase = ntile * bn, total_cols = p.group_cols;
        ok = mbar_wait(bar_tfull + 8 * as, aph, p.errflag, 3);
        if (!ok) break;
        tc_fence_after();
        {
          // plain convs with direct stores (res conv3: two 144-column tiles): this warp's chunks cg, cg + 4, ... (at most four)
          // come two per TMEM round trip, the accumulators are released before the last stores, the addressing is hoisted -
          // the general chunk loop below costs ~200 instructions per chunk and made conv3 epilogue-issue bound (ncu: 52 %
          // issue utilisation with the MMA warp waiting 43 % of the time for TMEM).
          const int phase = p.group_phase[0];
          const size_t pix1 = ((size_t)n * OH + gyb * out_scale + (phase >> 1)) * OW + gx * out_scale + (phase & 1);
          h16* const dst = (h16*)e.out + pix1 * e.out_ld + e.out_coff + cbase + cg * 16;
          const float* const b0 = bias_s + cbase + cg * 16;
          const int lim = e.out_c - cbase - 16;          // a chunk is live while its first column is <= lim
#pragma unroll 1
          for (int k0 = 0; k0 < 4; k0 += 2) {            // two chunks per TMEM round trip (four would spill)
            const int c0 = cg * 16 + 64 * k0, c1 = c0 + 64;
            const bool on0 = c0 < total_cols && c0 <= lim, on1 = c1 < total_cols && c1 <= lim;
            float v[2][16];
            if (on0) tmem_ld16_nowait(acc + (uint32_t)c0, v[0]);
            if (on1) tmem_ld16_nowait(acc + (uint32_t)c1, v[1]);
            tmem_ld_wait();
            if (k0 == 2) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(bar_tempty + 8 * as);
            }
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              if (!(k ? on1 : on0)) continue;
              const float* bk = b0 + 64 * (k0 + k);
              uint32_t o[8];
#pragma unroll
              for (int i = 0; i < 16; i += 4) {
                const float4 b4 = *reinterpret_cast<const float4*>(bk + i);
                float a0 = v[k][i] + b4.x, a1 = v[k][i + 1] + b4.y, a2 = v[k][i + 2] + b4.z, a3 = v[k][i + 3] + b4.w;
                if (e.act) {
                  a0 = fmaxf(a0, kLeaky * a0); a1 = fmaxf(a1, kLeaky * a1);
                  a2 = fmaxf(a2, kLeaky * a2); a3 = fmaxf(a3, kLeaky * a3);
                }
                o[i >> 1] = pack_h16x2(a0, a1);
                o[(i >> 1) + 1] = pack_h16x2(a2, a3);
              }
              if (!(p.ablate & 1))
                st_global_256(dst + 64 * (k0 + k), make_uint4(o[0], o[1], o[2], o[3]), make_uint4(o[4], o[5], o[6], o[7]));
            }
          }
        }
      } else if (EPI == EPI_HEADS) {
        // tile = R image rows of 256 pixels; accumulator group (half, row rr) = 16 columns [kw*2 + o] holding the
        // vertical 7x1 partial sums Y[x][kw][o];  out[x][o] = sum_kw Y[x+kw-3][kw][o] (zero outside the row)
        // + bias (model.py:246-247), then the grey composition (model.py:250-252).
        // 16 warps = 2 row slots x 2 halves x 4 lane quarters: two image rows are finished concurrently.
        const int half = cg & 1, slot = cg >> 1;
        const int y0 = tr * rows_per_tile;
        const int xg = half * 128 + r;
        // Horizontal taps by warp shuffle (see EPI_CLR): pixel x needs Y[x + kw - 3][kw]; only the three edge lanes on
        // either side of a warp go through shared memory.  edge[slot][row parity][warp in row][side][lane 0..2][kw 0..2][o]:
        // side 0 = lanes 29..31, groups kw = 0..2 (for the next warp), side 1 = lanes 0..2, groups kw = 4..6 (previous warp).
        const int wr = half * 4 + q;
        float* edge_slot = epi_work + slot * (2 * 8 * 40);
        const float b2 = bias_s[0], b3 = bias_s[1];
        const size_t p0 = ((size_t)n * OH + y0 + slot) * OW + xg;
        float i0 = x.img[3 * p0], i1 = x.img[3 * p0 + 1], i2 = x.img[3 * p0 + 2];     // prefetched one row ahead
        ok = mbar_wait(bar_tfull + 8 * as, aph, p.errflag, 3);
        if (!ok) break;
        tc_fence_after();
        for (int rr = slot; rr < rows_per_tile; rr += 2) {
          const size_t pidx = p0 + (size_t)(rr - slot) * OW;
          const float g = i0 * kGrayR + i1 * kGrayG + i2 * kGrayB;
          if (rr + 2 < rows_per_tile) {
            const size_t pn = pidx + 2 * (size_t)OW;
            i0 = x.img[3 * pn]; i1 = x.img[3 * pn + 1]; i2 = x.img[3 * pn + 2];
          }
          float v[16];
          tmem_ld16(acc + (uint32_t)((half * rows_per_tile + rr) * 16), v);
          if (p.zero_acc) tmem_zero16(acc + (uint32_t)((half * rows_per_tile + rr) * 16));
          if (rr + 2 >= rows_per_tile) {
            if (p.zero_acc) tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8 * as);        // accumulators drained: MMA may start the next tile
          }
          float* edge = edge_slot + (((rr - slot) >> 1) & 1) * (8 * 40);       // double-buffered by row parity: one barrier per row
          if (lane >= 29) {
#pragma unroll
            for (int i = 0; i < 6; ++i) edge[wr * 40 + (lane - 29) * 6 + i] = v[i];
          }
          if (lane < 3) {
#pragma unroll
            for (int i = 0; i < 6; ++i) edge[wr * 40 + 18 + lane * 6 + i] = v[8 + i];
          }
          epi_bar_slot(slot);
          float c2 = b2, c3 = b3;
#pragma unroll
          for (int kw = 0; kw < 7; ++kw) {
            const int src = lane + kw - 3;
            float t2 = __shfl_sync(0xffffffffu, v[2 * kw], src & 31), t3 = __shfl_sync(0xffffffffu, v[2 * kw + 1], src & 31);
            if (kw < 3 && src < 0) {                                  // previous warp's lane 32 + src = 29 + (src + 3)
              const float* e2 = edge + (wr - 1) * 40 + (src + 3) * 6 + 2 * kw;
              t2 = wr > 0 ? e2[0] : 0.f;
              t3 = wr > 0 ? e2[1] : 0.f;
            }
            if (kw > 3 && src > 31) {                                 // next warp's lane src - 32
              const float* e2 = edge + (wr + 1) * 40 + 18 + (src - 32) * 6 + 2 * (kw - 4);
              t2 = wr < 7 ? e2[0] : 0.f;
              t3 = wr < 7 ? e2[1] : 0.f;
            }
            c2 += t2;
            c3 += t3;
          }
          const float mask = tanhf(c2);
          const float gs = g * (1.f + mask) + c3;
          x.difgs[pidx] = gs - g;
          x.gs_f32[pidx] = gs;
          if (x.gs_out) x.gs_out[pidx] = gs;
          if (x.mask22_out) {
            x.mask22_out[3 * pidx] = fmaxf(mask, 0.f);
            x.mask22_out[3 * pidx + 1] = mask * 0.f;
            x.mask22_out[3 * pidx + 2] = fmaxf(-mask, 0.f);
          }
        }
      } else {
        // EPI_CLR: clr_conv1 over f (tensor cores, "kw expansion": group (half, row rr) = 48 columns [kw*16 + o] of
        // vertical 3x1 partial sums) + the gs channel of the concat (model.py:267) as a 3x3 fp32 conv on CUDA
        // cores, then clr_conv2 / clr_conv3 and the final dif (model.py:268-269, 288).
        // 16 warps = 2 row slots x 2 halves x 4 lane quarters (rows_per_tile = 2: one row per slot).
        const int half = cg & 1, slot = cg >> 1;
        const int y0 = tr * rows_per_tile;
        const int xg = half * 128 + r;
        // Horizontal taps: out[x] = Y[x-1][kw=0] + Y[x][kw=1] + Y[x+1][kw=2].  A thread owns pixel x, so the neighbours'
        // groups come by warp shuffle; only the two edge lanes of a warp go through shared memory (2 x 16 floats per
        // warp instead of 48 written + 48 read per THREAD: the row-exchange buffer used to compete with the UMMA operand
        // reads for the shared-memory pipe and slowed the MMAs to 2.4x their floor, profiles/r2_role_timers).
        // edge[slot][tile parity][warp in row 0..7][side][16]; the parity makes one barrier per tile sufficient.
        const int wr = half * 4 + q;
        float* edge = epi_work + ((slot * 2 + (int)(tcount & 1u)) * 8) * 32;
        const float* gsn = x.gs_f32 + (size_t)n * OH * OW;
        // issue the 9 gs taps and the input pixel first: their latency overlaps the accumulator wait
        float gv[9];
        const int gy = y0 + slot;
        const size_t pidx = ((size_t)n * OH + gy) * OW + xg;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const int yy = gy + t / 3 - 1, xx = xg + t % 3 - 1;
          gv[t] = (yy >= 0 && yy < OH && xx >= 0 && xx < OW) ? gsn[(size_t)yy * OW + xx] : 0.f;
        }
        const float i0 = x.img[3 * pidx], i1 = x.img[3 * pidx + 1], i2 = x.img[3 * pidx + 2];
        ok = mbar_wait(bar_tfull + 8 * as, aph, p.errflag, 3);
        if (!ok) break;
        tc_fence_after();
        float y0v[16], y1v[16], y2v[16];
        {
          const uint32_t ta = acc + (uint32_t)((half * rows_per_tile + slot) * CLR_NB);
          tmem_ld16_nowait(ta, y0v);
          tmem_ld16_nowait(ta + 16u, y1v);
          tmem_ld16_nowait(ta + 32u, y2v);
          tmem_ld_wait();
          if (p.zero_acc) {          // leave the group zeroed for the next tile's all-accumulate program
            tmem_zero16(ta);
            tmem_zero16(ta + 16u);
            tmem_zero16(ta + 32u);
            tmem_wait_st();
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tempty + 8 * as);
        if (lane == 31) {
#pragma unroll
          for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(edge + wr * 32 + i) = make_float4(y0v[i], y0v[i + 1], y0v[i + 2], y0v[i + 3]);
        }
        if (lane == 0) {
#pragma unroll
          for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(edge + wr * 32 + 16 + i) = make_float4(y2v[i], y2v[i + 1], y2v[i + 2], y2v[i + 3]);
        }
        epi_bar_slot(slot);
        float v[16];
        // the neighbour warps' edge groups and the bias come as 128-bit shared-memory loads (one wavefront per four values)
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f), u4 = t4;
          if (lane == 0 && wr > 0) t4 = *reinterpret_cast<const float4*>(edge + (wr - 1) * 32 + i);
          if (lane == 31 && wr < 7) u4 = *reinterpret_cast<const float4*>(edge + (wr + 1) * 32 + 16 + i);
          const float4 b4 = *reinterpret_cast<const float4*>(bias_s + i);
          const float el[4] = {t4.x, t4.y, t4.z, t4.w}, er[4] = {u4.x, u4.y, u4.z, u4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float l = __shfl_up_sync(0xffffffffu, y0v[i + k], 1);            // Y[x-1][kw=0]
            float rg = __shfl_down_sync(0xffffffffu, y2v[i + k], 1);         // Y[x+1][kw=2]
            if (lane == 0) l = el[k];
            if (lane == 31) rg = er[k];
            v[i + k] = ((bb[k] + l) + y1v[i + k]) + rg;
          }
        }
        // the colour tail is bound by its instruction count (profiles/r2_role_timers_mb256_after.txt): packed FFMA2
#pragma unroll
        for (int t = 0; t < 9; ++t) {
#pragma unroll
          for (int i = 0; i < 16; i += 2) ffma2(v[i], v[i + 1], gv[t], cw.wg[t * 16 + i], cw.wg[t * 16 + i + 1]);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], kLeaky * v[i]);
        float hbuf[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) hbuf[i] = cw.b2[i];
#pragma unroll
        for (int c = 0; c < 16; ++c) {
#pragma unroll
          for (int i = 0; i < 16; i += 2) ffma2(hbuf[i], hbuf[i + 1], v[c], cw.w2[c * 16 + i], cw.w2[c * 16 + i + 1]);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) hbuf[i] = fmaxf(hbuf[i], kLeaky * hbuf[i]);
        float rgb[3];
#pragma unroll
        for (int o = 0; o < 3; ++o) {
          // four independent partial sums keep the FMA chain short
          float s0 = cw.b3[o], s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            ffma2v(s0, s1, hbuf[4 * i], hbuf[4 * i + 1], cw.w3t[o * 16 + 4 * i], cw.w3t[o * 16 + 4 * i + 1]);
            ffma2v(s2, s3, hbuf[4 * i + 2], hbuf[4 * i + 3], cw.w3t[o * 16 + 4 * i + 2], cw.w3t[o * 16 + 4 * i + 3]);
          }
          rgb[o] = (s0 + s1) + (s2 + s3);
        }
        if (x.rgb_out) {
          x.rgb_out[3 * pidx] = rgb[0];
          x.rgb_out[3 * pidx + 1] = rgb[1];
          x.rgb_out[3 * pidx + 2] = rgb[2];
        }
        if (x.dif_out) {
          const float g1 = rgb[0] * kGrayR + rgb[1] * kGrayG + rgb[2] * kGrayB;
          const float g0 = i0 * kGrayR + i1 * kGrayG + i2 * kGrayB;
          x.dif_out[pidx] = g1 - g0;
        }
      }
    }
    if (tm) { p.timers[9] = BSR_CLK() - t_start; p.timers[10] = t_wtfull; p.timers[11] = tcount; }
    if (p.st_chunk && ethread == 0) bulk_wait0();        // staged stores have landed before the CTA retires
  }
  tc_fence_before();
  __syncthreads();
  if (warp == TC_EPI_WARPS + 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------
// host side
inline int configure_tc_kernels_conv() {
  cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<EPI_GENERIC, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_kernel<EPI_GENERIC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_kernel<EPI_HEADS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_kernel<EPI_CLR, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_kernel<EPI_PLAIN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_kernel<EPI_QKV, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  return e == cudaSuccess ? 0 : (int)e;
}

// host-side float -> 16-bit storage bits (round to nearest even; binary16 saturates to +-65504 like the device path)
inline uint16_t f32_to_h16_bits(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
#ifdef BSR_ACT_BF16
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
#else
  const uint16_t sign = (uint16_t)((u >> 16) & 0x8000u);
  const uint32_t a = u & 0x7fffffffu;
  if (a > 0x7f800000u) return (uint16_t)(sign | 0x7e00u);                  // NaN
  if (a >= 0x477ff000u) return (uint16_t)(sign | 0x7bffu);                 // >= 65520 (rounds past max) or inf: saturate
  if (a < 0x33000001u) return sign;                                        // <= 2^-25: rounds to zero
  const int e = (int)(a >> 23) - 127;
  uint32_t m = (a & 0x7fffffu) | 0x800000u;                                // 24-bit significand
  int shift = e >= -14 ? 13 : 13 + (-14 - e);                              // subnormal: shift further
  const uint32_t half = 1u << (shift - 1), rem = m & ((1u << shift) - 1u);
  uint32_t q = m >> shift;
  if (rem > half || (rem == half && (q & 1u))) ++q;
  const uint32_t bits = e >= -14 ? ((uint32_t)(e + 14) << 10) + q : q;     // carry out of q bumps the exponent correctly
  return (uint16_t)(sign | bits);
#endif
}

inline bool tc_upload(TmaEncoder& tma, TcWeights& t, const std::vector<uint16_t>& host, size_t rows, size_t K,
                      std::string* why) {
  if (cudaMalloc(&t.dev, host.size() * 2) != cudaSuccess) { *why = "cudaMalloc failed"; return false; }
  if (cudaMemcpy(t.dev, host.data(), host.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess) { *why = "cudaMemcpy failed"; return false; }
  uint64_t dims[2] = {K, rows}, strides[1] = {K * 2};
  uint32_t box[2] = {TC_BK, (uint32_t)t.b_box_rows};
  if (!tma.encode_h16(&t.map, t.dev, 2, dims, strides, box, nullptr)) { *why = tma.last_error; return false; }
  for (int i = 0; i < TC_MAX_STEPS; ++i) if ((t.steps[i].n_a & 7) == 0) t.steps[i].n_a |= 1;
  if (cudaMalloc(&t.steps_dev, sizeof t.steps) != cudaSuccess) { *why = "cudaMalloc failed"; return false; }
  if (cudaMemcpy(t.steps_dev, t.steps, sizeof t.steps, cudaMemcpyHostToDevice) != cudaSuccess) { *why = "cudaMemcpy failed"; return false; }
  t.b_total_rows = (int)rows;
  if (!t.b_resident && t.kind != TC_ROWPACK && t.kind != TC_HEADS && t.kind != TC_CLR && rows * K * 2 <= 150 * 1024 &&
      !getenv("BSR_NO_RESIDENT")) {
    // small weight matrices MAY stay resident in shared memory ([kblock][row] layout, no per-step weight TMA);
    // decided per launch: it only pays when a CTA processes many tiles
    t.can_reside = 1;
    t.b_res_kblocks = (int)(K / TC_BK);
  } else if (!t.b_resident && t.kind == TC_CONV && t.n_tiles > 1 && (size_t)t.bn * K * 2 <= 150 * 1024 &&
             !getenv("BSR_NO_RESIDENT") && !getenv("BSR_NO_PIN")) {
    t.can_pin = 1;
    t.b_res_kblocks = (int)(K / TC_BK);
  }
  t.ready = true;
  return true;
}

inline bool tc_disabled(const std::string& name) {
  const char* dis = getenv("BSR_TC_DISABLE");
  if (!dis) return false;
  std::string d = std::string(",") + dis + ",";
  return d.find("," + name + ",") != std::string::npos;
}

// Vertical-tap multi-row tiles: input row j (0 .. R+KH-2) of a tile feeds output row rr = j - kh with filter row kh.
// Weights are stored in DESCENDING filter-row order (block i <-> kh = KH-1-i) so that, for one input row, the
// weight blocks of consecutive output rows are consecutive too: all rows that already hold a partial sum are
// updated by ONE wide MMA (A read from shared memory once), and the row whose first tap this is by a second MMA
// with accumulate = 0.  Accumulator column of (group base, row rr) = (base + rr) * nb.
inline void tc_rows_sub(TcStep& sp, int j, int R, int KH, int nb, int base, int sub, bool merge = false) {
  const int lo = j - (KH - 1) > 0 ? j - (KH - 1) : 0, hi = j - 1 < R - 1 ? j - 1 : R - 1;
  const int16_t sb = (int16_t)(sub << 1);
  if (merge) {        // zero_acc: the row whose first tap this is (row j) accumulates like the others -> one MMA
    const int hi2 = j < R - 1 ? j : R - 1;
    if (hi2 >= lo)
      sp.mma[sp.n_mma++] = TcMma{(int16_t)((base + lo) * nb), (int16_t)((hi2 - lo + 1) * nb), (int16_t)((lo + KH - 1 - j) * nb), sb};
    return;
  }
  if (hi >= lo)
    sp.mma[sp.n_mma++] = TcMma{(int16_t)((base + lo) * nb), (int16_t)((hi - lo + 1) * nb), (int16_t)((lo + KH - 1 - j) * nb), sb};
  if (j <= R - 1) sp.mma[sp.n_mma++] = TcMma{(int16_t)((base + j) * nb), (int16_t)nb, (int16_t)((KH - 1) * nb), (int16_t)(sb | 1)};
}
// One step = input rows j and j+1 (two A sub-tiles, TMA boxes at dy and dy+1) when `pair` and both rows exist.
inline void tc_rows_step(TcStep& sp, int j, int R, int KH, int nb, int base, bool pair, bool merge = false) {
  sp.n_mma = 0;
  tc_rows_sub(sp, j, R, KH, nb, base, 0, merge);
  if (pair && j + 1 < R + KH - 1) {
    sp.n_a = 3;
    tc_rows_sub(sp, j + 1, R, KH, nb, base, 1, merge);
  }
}

// Build the step program and the packed h16 weight matrix of one layer from canonical fp32
// [tap][cin][cout].  Returns false with empty *why when the layer stays on the CUDA-core kernel.
inline bool pack_tc_weights(TmaEncoder& tma, const std::string& name, int kh, int kw, int cin, int cout, int transposed,
                            const std::vector<float>& w, const std::vector<float>& bias, TcWeights* out,
                            std::string* why) {
  why->clear();
  if (name == "clr_conv2" || name == "clr_conv3") return false;   // consumed by the fused colour tail
  if (tc_disabled(name)) return false;
  TcWeights& t = *out;
  memset(t.steps, 0, sizeof t.steps);
  t.kh = kh; t.kw = kw; t.cin = cin; t.cout = cout; t.transposed = transposed;
  t.tile_w = 0; t.a_sub = 1; t.rows_per_tile = 1; t.halves = 1; t.b_resident = 0; t.can_reside = 0; t.b_res_kblocks = 1;
  t.zero_acc = 0;
  auto W = [&](int tap, int c, int o) { return w[((size_t)tap * cin + c) * cout + o]; };

  if (name == "conv1") {
    // 7x7 conv over 3 channels (model.py:203).  The packed image (glue.cuh: pack_img_kernel) holds TWO image rows per
    // packed row, so one 64-element K block = 8 window pixels x [row y: 3 + 1 channels | row y + 1: 3 + 1 channels] covers
    // two filter rows (42 of 64 K elements are real, twice the density of one row per block).
    if (kh != 7 || kw != 7 || cin > 4 || transposed) { *why = "conv1 must be 7x7 with <= 4 input channels"; return false; }
    // Tile = R = 8 output rows x 128 pixels.  Input row pair p (tile rows 2p-3, 2p-2) feeds output row r with the filter
    // rows (k, k + 1), k = 2p - r in [-1, 6] (rows -1 and 7 are zero).  Weight blocks are stored in DESCENDING k, so the
    // output rows a pair updates are consecutive weight rows: ONE wide accumulate MMA for the rows that already hold a
    // partial sum plus one overwrite MMA for the (up to two) rows whose first pair this is.  32 KB of weights stay resident.
    const int R = 8, nb = (cout + 15) / 16 * 16;
    t.kind = TC_ROWPACK; t.cin_pad = 64; t.bn = R * nb; t.n_tiles = 1;
    t.b_box_rows = nb; t.b_stage_rows = 0; t.b_resident = 1; t.rows_per_tile = R; t.halves = 1;
    t.zero_acc = (R * nb <= 256 && !getenv("BSR_NO_ZERO_ACC")) ? 1 : 0;
    const size_t K = 64, rows = 8 * (size_t)nb;
    std::vector<uint16_t> host(rows * K, 0);
    for (int i = 0; i < 8; ++i) {                       // block i <-> k = 6 - i
      const int k = 6 - i;
      for (int b = 0; b < 7; ++b)
        for (int c = 0; c < cin; ++c)
          for (int o = 0; o < cout; ++o) {
            if (k >= 0) host[((size_t)i * nb + o) * K + b * 8 + c] = f32_to_h16_bits(W(k * 7 + b, c, o));
            if (k + 1 <= 6) host[((size_t)i * nb + o) * K + b * 8 + 4 + c] = f32_to_h16_bits(W((k + 1) * 7 + b, c, o));
          }
    }
    int ns = 0;
    for (int pr = 0; pr < 7; ++pr) {
      TcStep& sp = t.steps[ns++];
      // packed row yp = y + 1 holds image rows (y, y + 1); the pair starts at image row y0 - 3 + 2 pr
      sp.dy = (int8_t)(2 * pr - 3 + 1); sp.dx = 0; sp.a_c0 = 0; sp.b_rows = 0; sp.b_k = 0; sp.b_row = 0; sp.n_a = 1;
      sp.n_mma = 0;
      const int lo = 2 * pr - 6 > 0 ? 2 * pr - 6 : 0, hi = 2 * pr - 1 < R - 1 ? 2 * pr - 1 : R - 1;
      if (t.zero_acc) {      // accumulators start out zeroed: the pair's new rows ride in the same MMA (N <= 8 nb = 256)
        const int hi2 = 2 * pr + 1 < R - 1 ? 2 * pr + 1 : R - 1;
        sp.mma[sp.n_mma++] = TcMma{(int16_t)(lo * nb), (int16_t)((hi2 - lo + 1) * nb), (int16_t)((6 - 2 * pr + lo) * nb), 0};
        continue;
      }
      if (hi >= lo)
        sp.mma[sp.n_mma++] = TcMma{(int16_t)(lo * nb), (int16_t)((hi - lo + 1) * nb), (int16_t)((6 - 2 * pr + lo) * nb), 0};
      if (2 * pr <= R - 1) {
        const int cnt = 2 * pr + 1 <= R - 1 ? 2 : 1;
        sp.mma[sp.n_mma++] = TcMma{(int16_t)(2 * pr * nb), (int16_t)(cnt * nb), (int16_t)(6 * nb), 1};
      }
    }
    t.n_steps = ns;
    return tc_upload(tma, t, host, rows, K, why);
  }

  if (name == "heads") {
    // conv2|conv3 (7x7, 64 -> 2): accumulator column kw*2+o = sum_kh sum_c in[y+kh-3][x][c] * W[kh][kw][c][o]
    if (kh != 7 || kw != 7 || cin != 64 || cout != 2) { *why = "heads must be 7x7 64->2"; return false; }
    const int R = 8;
    t.kind = TC_HEADS; t.cin_pad = 64; t.bn = R * 2 * 16; t.n_tiles = 1; t.b_box_rows = 16; t.b_stage_rows = 0;
    t.b_resident = 1; t.rows_per_tile = R; t.halves = 2; t.tile_w = 128;
    t.zero_acc = getenv("BSR_NO_ZERO_ACC") ? 0 : 1;
    const size_t K = 64, rows = 7 * 16;
    std::vector<uint16_t> host(rows * K, 0);
    for (int a = 0; a < 7; ++a)
      for (int b = 0; b < 7; ++b)
        for (int c = 0; c < 64; ++c)
          for (int o = 0; o < 2; ++o) host[((size_t)(6 - a) * 16 + b * 2 + o) * K + c] = f32_to_h16_bits(W(a * 7 + b, c, o));
    int ns = 0;
    const bool rpair = !getenv("BSR_NO_RPAIR");
    if (rpair) t.a_sub = 2;
    for (int hh = 0; hh < 2; ++hh)
      for (int j = 0; j < R + 6; j += rpair ? 2 : 1) {
        TcStep& sp = t.steps[ns++];
        sp.dy = (int8_t)(j - 3); sp.dx = 0; sp.a_x0 = (int16_t)(hh * 128); sp.a_c0 = 0;
        tc_rows_step(sp, j, R, 7, 16, hh * R, rpair, t.zero_acc != 0);
      }
    t.n_steps = ns;
    return tc_upload(tma, t, host, rows, K, why);
  }

  if (name == "clr_conv1") {
    // canonical input order [f0..f63, gs]: f runs on tensor cores, gs (fp32) in the epilogue
    if (kh != 3 || kw != 3 || cin != 65 || cout != 16) { *why = "clr_conv1 must be 3x3 65->16"; return false; }
    const int R = 2, nb = CLR_NB;
    t.kind = TC_CLR; t.cin_pad = 64; t.bn = R * 2 * nb; t.n_tiles = 1; t.b_box_rows = nb; t.b_stage_rows = 0;
    t.b_resident = 1; t.rows_per_tile = R; t.halves = 2; t.tile_w = 128;
    t.zero_acc = getenv("BSR_NO_ZERO_ACC") ? 0 : 1;          // 4 MMA groups per half tile instead of 5 (see heads)
    const size_t K = 64, rows = 3 * (size_t)nb;
    std::vector<uint16_t> host(rows * K, 0);
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b)
        for (int c = 0; c < 64; ++c)
          for (int o = 0; o < 16; ++o) host[((size_t)(2 - a) * nb + b * 16 + o) * K + c] = f32_to_h16_bits(W(a * 3 + b, c, o));
    int ns = 0;
    const bool rpair = !getenv("BSR_NO_RPAIR");
    if (rpair) t.a_sub = 2;
    for (int hh = 0; hh < 2; ++hh)
      for (int j = 0; j < R + 2; j += rpair ? 2 : 1) {
        TcStep& sp = t.steps[ns++];
        sp.dy = (int8_t)(j - 1); sp.dx = 0; sp.a_x0 = (int16_t)(hh * 128); sp.a_c0 = 0;
        tc_rows_step(sp, j, R, 3, nb, hh * R, rpair, t.zero_acc != 0);
      }
    t.n_steps = ns;
    (void)bias;
    return tc_upload(tma, t, host, rows, K, why);
  }

  t.cin_pad = (cin + 63) / 64 * 64;
  const int ncb = t.cin_pad / 64;
  // whole 16-element K steps at the end of the last K block that hold only padding: not issued (TcStep.n_a bits 4-5)
  const int kskip = getenv("BSR_NO_KSKIP") ? 0 : (t.cin_pad - cin) / 16;
  if (transposed && cout % 16 == 0 && cout <= 96 && 4 * ncb <= TC_MAX_STEPS && !tc_disabled("convt_fused")) {
    // fused 4-phase transposed conv
    if (kh != 3 || kw != 3) { *why = "transposed conv must be 3x3"; return false; }
    t.kind = TC_CONVT_FUSED; t.bn = 4 * cout; t.n_tiles = 1; t.b_box_rows = cout; t.b_stage_rows = 4 * cout;
    const size_t K = t.cin_pad, rows = 9 * (size_t)cout;
    std::vector<uint16_t> host(rows * K, 0);
    // Accumulator column groups are [p10 | p00 | p01 | p11] so that the two phases fed by each one-pixel shift are
    // adjacent and take ONE MMA of N = 2*cout (an SS-mode MMA costs ~44 + 0.47 N cycles: fewer, wider is cheaper).
    // weight row blocks, in column order: shift(0,0): taps (1,0),(0,0),(0,1),(1,1); shift(0,-1): (1,2),(0,2);
    // shift(-1,0): (2,0),(2,1); shift(-1,-1): (2,2)
    const int order[9][2] = {{1, 0}, {0, 0}, {0, 1}, {1, 1}, {1, 2}, {0, 2}, {2, 0}, {2, 1}, {2, 2}};
    for (int b = 0; b < 9; ++b) {
      const int tap = order[b][0] * 3 + order[b][1];
      for (int c = 0; c < cin; ++c)
        for (int o = 0; o < cout; ++o) host[((size_t)b * cout + o) * K + c] = f32_to_h16_bits(W(tap, c, o));
    }
    int ns = 0;
    const int16_t co = (int16_t)cout;
    // two K blocks per step only if two such stages still fit the shared-memory budget when weights stream
    const bool pair_k = !getenv("BSR_NO_KPAIR") && 2 * 2 * (TC_BM * 128 + 4 * cout * 128) <= TC_SMEM_BUDGET;
    for (int cb = 0; cb < ncb;) {
      const int na = (pair_k && ncb - cb >= 2) ? 2 : 1;
      const int16_t f = (int16_t)(cb == 0);
      TcStep s;
      memset(&s, 0, sizeof s);
      s.a_c0 = (int16_t)(cb * 64); s.b_k = cb * 64; s.n_a = (int8_t)(na | (cb + na == ncb ? kskip << 4 : 0));
      if (na == 2) t.a_sub = 2;
      // shift (0,0): all four phases
      s.dy = 0; s.dx = 0; s.b_row = 0; s.b_rows = (int16_t)(4 * co);
      if (4 * cout <= 256) { s.n_mma = 1; s.mma[0] = TcMma{0, (int16_t)(4 * co), 0, f}; }
      else { s.n_mma = 2; s.mma[0] = TcMma{0, (int16_t)(2 * co), 0, f}; s.mma[1] = TcMma{(int16_t)(2 * co), (int16_t)(2 * co), (int16_t)(2 * co), f}; }
      t.steps[ns++] = s;
      // shift (0,-1): kw = 2 -> phases p10, p00 (columns 0 .. 2*cout)
      s.dy = 0; s.dx = -1; s.b_row = (int16_t)(4 * co); s.b_rows = (int16_t)(2 * co);
      s.n_mma = 1; s.mma[0] = TcMma{0, (int16_t)(2 * co), 0, 0};
      t.steps[ns++] = s;
      // shift (-1,0): kh = 2 -> phases p00, p01 (columns cout .. 3*cout)
      s.dy = -1; s.dx = 0; s.b_row = (int16_t)(6 * co); s.b_rows = (int16_t)(2 * co);
      s.n_mma = 1; s.mma[0] = TcMma{co, (int16_t)(2 * co), 0, 0};
      t.steps[ns++] = s;
      // shift (-1,-1): tap (2,2) -> p00 (columns cout .. 2*cout)
      s.dy = -1; s.dx = -1; s.b_row = (int16_t)(8 * co); s.b_rows = co;
      s.n_mma = 1; s.mma[0] = TcMma{co, co, 0, 0};
      t.steps[ns++] = s;
      cb += na;
    }
    t.n_steps = ns;
    // Halo program (used when a tile is one image row of 128 pixels and the weights are resident): the four shifted A
    // operands (0,0), (0,-1), (-1,0), (-1,-1) are two 129-pixel row tiles fetched at x0 - 1; the x-shift is a view that
    // starts one pixel row later.  2 steps of 2 K blocks per tile: half the L2 -> shared-memory traffic, and two stages
    // hold a whole tile instead of half of one.
    t.n_steps_halo = 0;
    if (ncb == 2 && pair_k && 4 * cout <= 256 && !getenv("BSR_NO_HALO")) {
      memset(t.steps_halo, 0, sizeof t.steps_halo);
      for (int r = 0; r < 2; ++r) {
        TcStep& h = t.steps_halo[r];
        h.dy = (int8_t)(-r); h.dx = -1; h.a_c0 = 0; h.b_k = 0; h.n_a = (int8_t)(2 | (kskip << 4)); h.n_mma = 2;
        if (r == 0) {
          h.b_row = 0; h.b_rows = (int16_t)(6 * co);
          h.mma[0] = TcMma{0, (int16_t)(4 * co), 0, (int16_t)(1 | 4)};                 // shift (0,0): all four phases, view + 1
          h.mma[1] = TcMma{0, (int16_t)(2 * co), (int16_t)(4 * co), 0};                // shift (0,-1): p10, p00
        } else {
          h.b_row = (int16_t)(6 * co); h.b_rows = (int16_t)(3 * co);
          h.mma[0] = TcMma{co, (int16_t)(2 * co), 0, 4};                                // shift (-1,0): p00, p01, view + 1
          h.mma[1] = TcMma{co, co, (int16_t)(2 * co), 0};                               // shift (-1,-1): p00
        }
      }
      t.n_steps_halo = 2;
    }
    if (!tc_upload(tma, t, host, rows, K, why)) return false;
    if (t.n_steps_halo) {
      if (cudaMalloc(&t.steps_halo_dev, sizeof t.steps_halo) != cudaSuccess ||
          cudaMemcpy(t.steps_halo_dev, t.steps_halo, sizeof t.steps_halo, cudaMemcpyHostToDevice) != cudaSuccess) {
        *why = "cudaMalloc / cudaMemcpy of the halo step program failed";
        return false;
      }
    }
    return true;
  }

  if (!transposed && kh == 3 && kw == 3 && cin == 32 && name == "down1" && !getenv("BSR_NO_PAIRX")) {
    // 3x3 / stride 2 over a 32-channel tensor (down1, model.py:207): two neighbouring pixels are 64 contiguous
    // elements, so taps (kh,0),(kh,1) of output column ox are ONE full K block at pixel pair ox and tap (kh,2) is the
    // first half of pair ox+1: 6 steps instead of 9 half-empty ones, and no x-stride in the TMA box.
    t.kind = TC_PAIRX; t.cin_pad = 64; t.bn = (cout + 15) / 16 * 16; t.n_tiles = 1;
    t.b_box_rows = t.bn; t.b_stage_rows = t.bn;
    const size_t K = 6 * 64, rows = t.bn;
    std::vector<uint16_t> host(rows * K, 0);
    // one step per filter row: its two K blocks (pixel pair ox and the first half of pair ox + 1) are x-neighbouring rows of
    // ONE 129-row halo tile (n_a = 5: the second block is a view one row in): 3 steps of 8 MMAs per tile instead of 6 of 4 -
    // the per-step control cost of the issuer (~400 cycles, profiles/r2_role_timers) is paid half as often - and 3 TMA
    // boxes per tile instead of 6
    const bool pair_x = !getenv("BSR_NO_KPAIR");
    int ns = 0, kb = 0;
    for (int a = 0; a < 3; ++a)
      for (int half = 0; half < 2; ++half, ++kb) {
        for (int o = 0; o < cout; ++o)
          for (int c = 0; c < 32; ++c) {
            if (half == 0) {
              host[(size_t)o * K + kb * 64 + c] = f32_to_h16_bits(W(a * 3 + 0, c, o));
              host[(size_t)o * K + kb * 64 + 32 + c] = f32_to_h16_bits(W(a * 3 + 1, c, o));
            } else {
              host[(size_t)o * K + kb * 64 + c] = f32_to_h16_bits(W(a * 3 + 2, c, o));
            }
          }
        if (pair_x && half == 1) continue;          // covered by the paired step of half 0
        TcStep& sp = t.steps[ns];
        sp.dy = (int8_t)a; sp.dx = (int8_t)half; sp.a_c0 = 0; sp.b_rows = (int16_t)t.bn; sp.b_k = kb * 64; sp.b_row = 0;
        sp.n_a = (int8_t)(pair_x ? 5 : 1);
        sp.n_mma = 1; sp.mma[0] = TcMma{0, (int16_t)t.bn, 0, (int16_t)(ns == 0)};
        ++ns;
      }
    t.n_steps = ns;                      // a_sub stays 1: the paired K block is a view of the same (129-row) tile
    return tc_upload(tma, t, host, rows, K, why);
  }

  // ---- plain conv, or one launch per phase for wide transposed convs (handled by the caller via `phase`)
  if (transposed) return false;          // cout > 96 transposed convs: see pack_tc_weights_phase
  if (kh * kw * ncb > TC_MAX_STEPS) return false;
  t.kind = TC_CONV;
  const bool pair_k = !getenv("BSR_NO_KPAIR");
  // q | k | v projection: two 192-column tiles (N = 192 MMAs read the A operand twice per pixel tile instead of three times
  // and amortise the per-MMA cost better than N = 128; measured below); BSR_QKV_128=1 restores three 128-column tiles
  if (cout == 384) { if (getenv("BSR_QKV_128")) { t.bn = 128; t.n_tiles = 3; } else { t.bn = 192; t.n_tiles = 2; } }
  else if (cout <= 256) { t.bn = (cout + 15) / 16 * 16; t.n_tiles = 1; }
  else { t.n_tiles = (cout + 255) / 256; t.bn = ((cout + t.n_tiles - 1) / t.n_tiles + 15) / 16 * 16; }
  t.b_box_rows = t.bn; t.b_stage_rows = t.bn;
  const size_t K = (size_t)kh * kw * t.cin_pad, rows = (size_t)t.n_tiles * t.bn;
  std::vector<uint16_t> host(rows * K, 0);
  int ns = 0;
  for (int tap = 0; tap < kh * kw; ++tap) {
    for (int c = 0; c < cin; ++c)
      for (int o = 0; o < cout; ++o) host[(size_t)o * K + (size_t)tap * t.cin_pad + c] = f32_to_h16_bits(W(tap, c, o));
    for (int cb = 0; cb < ncb;) {
      const int na = (pair_k && ncb - cb >= 2) ? 2 : 1;
      TcStep& s = t.steps[ns];
      s.dy = (int8_t)(tap / kw); s.dx = (int8_t)(tap % kw);        // SAME-padding offset subtracted at launch
      s.a_c0 = (int16_t)(cb * 64); s.b_rows = (int16_t)t.bn; s.b_k = (tap * ncb + cb) * 64; s.b_row = 0;
      s.n_a = (int8_t)(na | (cb + na == ncb ? kskip << 4 : 0));
      s.n_mma = 1; s.mma[0] = TcMma{0, (int16_t)t.bn, 0, (int16_t)(ns == 0)};
      if (na == 2) t.a_sub = 2;
      ++ns;
      cb += na;
    }
  }
  t.n_steps = ns;
  return tc_upload(tma, t, host, rows, K, why);
}

// Wide transposed convs (cout > 96: clr_up1) run as 4 launches, one per sub-pixel phase, each a plain
// gather conv over that phase's taps.
inline bool pack_tc_weights_phase(TmaEncoder& tma, int phase, int cin, int cout, const std::vector<float>& w,
                                  TcWeights* out, std::string* why) {
  TcWeights& t = *out;
  memset(t.steps, 0, sizeof t.steps);
  t.a_sub = 1; t.rows_per_tile = 1; t.halves = 1; t.b_resident = 0; t.can_reside = 0; t.b_res_kblocks = 1; t.tile_w = 0;
  t.kind = TC_CONV; t.kh = 3; t.kw = 3; t.cin = cin; t.cout = cout; t.transposed = 1;
  t.cin_pad = (cin + 63) / 64 * 64;
  const int ncb = t.cin_pad / 64, py = phase >> 1, px = phase & 1;
  const int kskip = getenv("BSR_NO_KSKIP") ? 0 : (t.cin_pad - cin) / 16;          // see pack_tc_weights
  t.bn = (cout + 15) / 16 * 16; t.n_tiles = 1; t.b_box_rows = t.bn; t.b_stage_rows = t.bn;
  std::vector<int> taps;
  for (int a = py; a < 3; a += 2)
    for (int b = px; b < 3; b += 2) taps.push_back(a * 3 + b);
  if ((int)taps.size() * ncb > TC_MAX_STEPS) { *why = "too many steps"; return false; }
  const size_t K = taps.size() * (size_t)t.cin_pad, rows = t.bn;
  std::vector<uint16_t> host(rows * K, 0);
  int ns = 0;
  for (size_t ti = 0; ti < taps.size(); ++ti) {
    const int tap = taps[ti];
    for (int c = 0; c < cin; ++c)
      for (int o = 0; o < cout; ++o)
        host[(size_t)o * K + ti * t.cin_pad + c] = f32_to_h16_bits(w[((size_t)tap * cin + c) * cout + o]);
    for (int cb = 0; cb < ncb;) {
      const int na = (!getenv("BSR_NO_KPAIR") && ncb - cb >= 2) ? 2 : 1;
      TcStep& s = t.steps[ns];
      s.dy = (int8_t)(-((tap / 3) >> 1)); s.dx = (int8_t)(-((tap % 3) >> 1));
      s.a_c0 = (int16_t)(cb * 64); s.b_rows = (int16_t)t.bn; s.b_k = (int)((ti * ncb + cb) * 64); s.b_row = 0;
      s.n_a = (int8_t)(na | (cb + na == ncb ? kskip << 4 : 0));
      s.n_mma = 1; s.mma[0] = TcMma{0, (int16_t)t.bn, 0, (int16_t)(ns == 0)};
      if (na == 2) t.a_sub = 2;
      ++ns;
      cb += na;
    }
  }
  t.n_steps = ns;
  return tc_upload(tma, t, host, rows, K, why);
}

struct TmapKey {
  const void* p; int ld, coff, cin, H, W, n, stride, bw, bh, kind;
  bool operator<(const TmapKey& o) const {
    return std::tie(p, ld, coff, cin, H, W, n, stride, bw, bh, kind) <
           std::tie(o.p, o.ld, o.coff, o.cin, o.H, o.W, o.n, o.stride, o.bw, o.bh, o.kind);
  }
};

// phase: -1 = whole layer; 0..3 = single sub-pixel phase of a wide transposed conv (weights from
// pack_tc_weights_phase).
inline int launch_conv_tc(TmaEncoder& tma, const TcWeights& t, const void* in, int in_ld, int in_coff, int H, int W,
                          int stride, int n, const EpiParams& e, const EpiExtra& x, int phase, int num_sms,
                          int* errflag, cudaStream_t st, int* launches, const Knobs& kn, PlanCounters* pc) {
  static thread_local std::map<TmapKey, CUtensorMap> cache;
  static thread_local ConvTcParams p;           // 2 KB: keep it off the stack
  memset(&p, 0, sizeof p);
  int GH, GW;
  int pad_t = 0, pad_l = 0;
  if (t.transposed) {
    GH = H; GW = W; p.in_stride = 1; p.out_scale = 2; p.OH = 2 * H; p.OW = 2 * W;
  } else {
    p.in_stride = stride; p.out_scale = 1;
    p.OH = (H + stride - 1) / stride; p.OW = (W + stride - 1) / stride;
    GH = p.OH; GW = p.OW;
    int tot_h = (p.OH - 1) * stride + t.kh - H; if (tot_h < 0) tot_h = 0;
    int tot_w = (p.OW - 1) * stride + t.kw - W; if (tot_w < 0) tot_w = 0;
    pad_t = tot_h / 2; pad_l = tot_w / 2;        // TF SAME: before = total // 2
  }
  p.in_stride_x = p.in_stride;
  if (t.kind == TC_PAIRX) {
    if (stride != 2 || (W & 1) || (H & 1) || in_ld != 32) { tma.last_error = "pair-packed conv needs stride 2 and a dense 32-channel input"; return -7; }
    p.in_stride_x = 1;           // TMA x dimension counts pixel pairs; SAME padding of a stride-2 conv on even sizes is 0 before
  }
  p.n_steps = t.n_steps;
  p.steps = t.steps_dev;
  if (t.kind == TC_CONV && !t.transposed) { p.pad_t = pad_t; p.pad_l = pad_l; }
  p.rows_per_tile = t.rows_per_tile; p.halves = t.halves;
  p.b_total_rows = t.b_total_rows;
  p.bw = t.tile_w ? t.tile_w : (GW < TC_BM ? GW : TC_BM);
  p.bh = TC_BM / p.bw;
  p.epi_mode = t.kind == TC_HEADS ? EPI_HEADS : (t.kind == TC_CLR ? EPI_CLR : EPI_GENERIC);
  if (t.halves == 2) {
    if (GW != 256 || GH % t.rows_per_tile) { tma.last_error = "row-tile kernels need 256-pixel rows"; return -4; }
    p.tiles_x = 1; p.tiles_y = GH / t.rows_per_tile;
  } else if (t.rows_per_tile > 1) {
    if (GW % p.bw || GH % t.rows_per_tile || p.bh != 1) { tma.last_error = "multi-row tiles need full 128-pixel rows"; return -4; }
    p.tiles_x = GW / p.bw; p.tiles_y = GH / t.rows_per_tile;
  } else {
    if (GW % p.bw || GH % p.bh) { tma.last_error = "spatial size not tileable into 128-pixel rectangles"; return -1; }
    p.tiles_x = GW / p.bw; p.tiles_y = GH / p.bh;
  }
  p.n_img = n; p.bn = t.bn; p.n_tiles = t.n_tiles; p.b_box_rows = t.b_box_rows;
  if (t.kind == TC_CONVT_FUSED) {
    p.n_groups = 4; p.group_cols = t.cout;
    p.group_phase[0] = 2; p.group_phase[1] = 0; p.group_phase[2] = 1; p.group_phase[3] = 3;   // p10 p00 p01 p11
  } else if (t.rows_per_tile > 1) {
    p.n_groups = t.rows_per_tile; p.group_cols = t.bn / t.rows_per_tile;      // TC_ROWPACK: one group per output row
  } else {
    p.n_groups = 1; p.group_cols = t.bn;
    p.group_phase[0] = phase < 0 ? 0 : phase;
  }
  p.total_tiles = n * p.tiles_x * p.tiles_y * t.n_tiles;
  int grid = p.total_tiles < num_sms ? p.total_tiles : num_sms;
  const bool pinned = t.can_pin && p.total_tiles >= 6 * num_sms;
  if (pinned) grid = num_sms / t.n_tiles * t.n_tiles;        // tile += grid keeps tile % n_tiles fixed per CTA
  const bool resident = t.b_resident || pinned || (t.can_reside && p.total_tiles >= 6 * grid);
  p.b_resident = resident ? 1 : 0;
  p.b_pinned = pinned ? 1 : 0;
  p.b_res_rows = pinned ? t.bn : t.b_total_rows;
  // bias staging: 512 floats unless the layer provably reads fewer (frees shared memory for the store staging)
  const int bias_floats = (p.epi_mode == EPI_GENERIC && p.n_groups > 1 && p.group_cols <= 128) ? 128 : 512;
  const int epi_bytes = bias_floats * 4 + (p.epi_mode == EPI_HEADS ? 2 * 2 * 8 * 40 * 4 : (p.epi_mode == EPI_CLR ? 2 * 2 * 8 * 32 * 4 : 0));
  p.epi_bytes = (epi_bytes + 127) / 128 * 128;
  p.steps_bytes = (int)((t.n_steps * sizeof(TcStep) + 127) / 128 * 128);
  p.b_res_kblocks = t.b_res_kblocks;
  // halo tiles (fused transposed conv, one 128-pixel image row per tile, resident weights): see pack_tc_weights
  const bool halo_ct = t.kind == TC_CONVT_FUSED && t.n_steps_halo > 0 && resident && p.bw == TC_BM && p.bh == 1 && !kn.no_halo;
  const bool halo = halo_ct || (t.kind == TC_PAIRX && t.steps[0].n_a == 5);
  if (t.kind == TC_PAIRX && t.steps[0].n_a == 5 && (p.bw != TC_BM || p.bh != 1)) { tma.last_error = "pair-packed halo tiles need 128-pixel rows"; return -8; }
  if (halo_ct) {
    p.n_steps = t.n_steps_halo;
    p.steps = t.steps_halo_dev;
    p.steps_bytes = (int)((t.n_steps_halo * sizeof(TcStep) + 127) / 128 * 128);
  }
  p.a_tile_bytes = halo ? 17 * 1024 : TC_BM * 128;
  p.a_box_bytes = halo ? (TC_BM + 1) * 128 : TC_BM * 128;
  p.ablate = kn.ablate;
  p.timers = reinterpret_cast<long long*>(errflag) + 16 + 16 * ((*launches) & 63);
  const int fixed_bytes = 1024 + (resident ? t.b_res_kblocks * p.b_res_rows * 128 : 0) + 192 + p.steps_bytes + p.epi_bytes + 512 + 64;
  p.a_sub = t.a_sub;
  const int b_sub = (t.kind == TC_PAIRX && t.steps[0].n_a == 5) ? 2 : t.a_sub;          // K blocks of weights per step
  p.b_kb_rows = t.b_stage_rows;
  p.zero_acc = t.zero_acc;
  p.stage_bytes = t.a_sub * p.a_tile_bytes + (resident ? 0 : b_sub * t.b_stage_rows * 128);
  const int max_smem = 227 * 1024;
  // Staged TMA-store epilogue (h16 NHWC outputs whose column groups are multiples of 64 channels = whole 64-column
  // iterations), used when the staging fits without shortening the TMA->MMA pipeline; otherwise the direct-store epilogue
  // runs.  32- and 96-channel groups (conv1, up1, clr_up2) are faster with direct 256-bit stores: measured on one box,
  // staged -> direct: conv1 0.535 -> 0.472 ms, clr_up2 0.543 -> 0.487, up1 0.253 -> 0.229 (256 images per launch).
  p.st_chunk = 0;
  int staging = 0;
  const int stages_direct = std::min(8, (max_smem - fixed_bytes) / p.stage_bytes);
  const bool st_ok = p.epi_mode == EPI_GENERIC && e.res1 == nullptr && e.res2 == nullptr && e.out_mode == OUT_T &&
                     t.n_tiles == 1 && p.group_cols % 64 == 0 && e.out_c == p.group_cols && e.out_ld % 8 == 0 &&
                     e.out_coff % 8 == 0 && p.OH % p.out_scale == 0 && p.OW % p.out_scale == 0 && !kn.no_tma_store;
  if (st_ok) {
    // 64 accumulator columns per iteration; double-buffered (2 x 16 KB) if the pipeline keeps its depth, else one buffer
    for (int bufs = 2; bufs >= 1 && !p.st_chunk; --bufs) {
      const int need = bufs * 128 * 64 * 2;
      const int ns = std::min(8, (max_smem - fixed_bytes - need) / p.stage_bytes);
      // one buffer costs a second barrier per iteration: only worth it where the pipeline has >= 3 stages of slack
      // (measured: up2 -13 %, conv1 -10 %, but up3 / clr_up3 with 2 stages +14 %)
      if (ns >= 2 && (ns >= stages_direct || ns >= 4) && (bufs == 2 || ns >= 3)) { p.st_chunk = 16; p.st_bufs = bufs; staging = need; }
    }
    if (kn.st_bufs) {      // experiment knob: force the buffer count where it fits
      const int bufs = kn.st_bufs, need = bufs * 128 * 64 * 2;
      if ((bufs == 1 || bufs == 2) && (max_smem - fixed_bytes - need) / p.stage_bytes >= 2) { p.st_chunk = 16; p.st_bufs = bufs; staging = need; }
    }
  }
  // lean direct-store epilogue (its own instantiation) for plain convs the staged epilogue does not take: res conv3
  if (p.epi_mode == EPI_GENERIC && !p.st_chunk && e.res1 == nullptr && e.res2 == nullptr && e.out_mode == OUT_T &&
      p.n_groups == 1 && t.rows_per_tile == 1 && !t.zero_acc && p.group_cols <= 256 && p.group_cols != 128 &&
      e.out_ld % 16 == 0 && e.out_coff % 16 == 0 && e.out_c % 16 == 0 && !getenv("BSR_NO_PLAIN_EPI"))
    p.epi_mode = EPI_PLAIN;
  if (p.epi_mode == EPI_GENERIC && e.out_mode == OUT_QKV && e.v_natural && e.res1 == nullptr && e.res2 == nullptr &&
      p.n_groups == 1 && t.rows_per_tile == 1 && p.group_cols == 192 && e.out_c == t.n_tiles * 192 && p.out_scale == 1) {
    p.epi_mode = EPI_QKV;
  }
  p.n_stages = std::min(8, (max_smem - fixed_bytes - staging) / p.stage_bytes);
  if (p.n_stages < 2) { tma.last_error = "stage too large"; return -5; }
  p.acc_stages = 2 * t.bn <= 512 ? 2 : 1;
  p.errflag = errflag;
  if ((in_ld % 8) || (in_coff % 8)) { tma.last_error = "input channel stride/offset must be multiples of 8"; return -2; }
  TmapKey key{in, in_ld, in_coff, t.cin, H, W, n, p.in_stride, p.bw, p.bh, t.kind + (halo ? 2000 : 0)};
  auto it = cache.find(key);
  if (it == cache.end()) {
    CUtensorMap m;
    uint64_t dims[4] = {(uint64_t)(t.kind == TC_CLR ? 64 : t.cin), (uint64_t)W, (uint64_t)H, (uint64_t)n};
    uint64_t strides[3] = {(uint64_t)in_ld * 2, (uint64_t)W * in_ld * 2, (uint64_t)H * W * in_ld * 2};
    if (t.kind == TC_ROWPACK) {
      // packed image [N][H + 1][W + 8][8] (two image rows per packed row): window of output column x = 64 contiguous
      // elements starting at packed pixel x (= image pixel x-3); consecutive windows overlap (dim-1 stride 16 B < dim-0
      // extent 128 B)
      dims[0] = 64; dims[2] = (uint64_t)(H + 1);
      strides[0] = 16; strides[1] = (uint64_t)(W + 8) * 16; strides[2] = (uint64_t)(H + 1) * (W + 8) * 16;
    }
    uint32_t box[4] = {TC_BK, (uint32_t)(p.bw * p.in_stride + (halo ? 1 : 0)), (uint32_t)(p.bh * p.in_stride), 1};
    uint32_t es[4] = {1, (uint32_t)p.in_stride, (uint32_t)p.in_stride, 1};
    if (t.kind == TC_PAIRX) {
      // view [N][H][W/2][64]: dim 0 = one pixel pair (2 x 32 channels, 128 contiguous bytes)
      dims[0] = 64; dims[1] = (uint64_t)(W / 2);
      strides[0] = 128;
      box[1] = (uint32_t)p.bw + (halo ? 1u : 0u); es[1] = 1;
    }
    if (!tma.encode_h16(&m, (void*)((const h16*)in + in_coff), 4, dims, strides, box, es)) return -3;
    if (cache.size() > 4096) cache.clear();
    it = cache.emplace(key, m).first;
  }
  // output map of the staged epilogue: [C, px, W/s, py, N*H/s] (s = out_scale; a 3x3/s2 transposed conv writes the
  // four sub-pixel phases (py, px) of every input position), box = 32 channels x bw x bh positions
  static const CUtensorMap no_map = {};
  const CUtensorMap* omap = &no_map;
  if (p.st_chunk) {
    const int sc = p.out_scale;
    TmapKey okey{e.out, e.out_ld, 0, 0, p.OH, p.OW, n, sc, p.bw, p.bh, 1000};
    auto ot = cache.find(okey);
    if (ot == cache.end()) {
      CUtensorMap m;
      const uint64_t ld2 = (uint64_t)e.out_ld * 2;
      uint64_t dims[5] = {(uint64_t)e.out_ld, (uint64_t)sc, (uint64_t)(p.OW / sc), (uint64_t)sc, (uint64_t)n * (uint64_t)(p.OH / sc)};
      uint64_t strides[4] = {ld2, (uint64_t)sc * ld2, (uint64_t)p.OW * ld2, (uint64_t)sc * p.OW * ld2};
      uint32_t box[5] = {32, 1, (uint32_t)p.bw, 1, (uint32_t)p.bh};
      if (!tma.encode_h16_store(&m, e.out, 5, dims, strides, box)) return -3;
      ot = cache.emplace(okey, m).first;       // std::map: `it` (the input map) stays valid
    }
    omap = &ot->second;
  }
  const size_t smem = (size_t)fixed_bytes + (size_t)staging + (size_t)p.n_stages * p.stage_bytes;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = kn.no_pdl ? 0 : 1;
  static const ClrWeights no_clr = {};
  const ClrWeights& cw = (p.epi_mode == EPI_CLR && t.clr) ? *t.clr : no_clr;
  cudaError_t le;
  if (p.epi_mode == EPI_HEADS) le = cudaLaunchKernelEx(&cfg, conv_tc_kernel<EPI_HEADS, false>, it->second, t.map, p, e, x, cw, *omap);
  else if (p.epi_mode == EPI_CLR) le = cudaLaunchKernelEx(&cfg, conv_tc_kernel<EPI_CLR, false>, it->second, t.map, p, e, x, cw, *omap);
  else if (p.epi_mode == EPI_PLAIN) le = cudaLaunchKernelEx(&cfg, conv_tc_kernel<EPI_PLAIN, false>, it->second, t.map, p, e, x, cw, *omap);
  else if (p.epi_mode == EPI_QKV) le = cudaLaunchKernelEx(&cfg, conv_tc_kernel<EPI_QKV, false>, it->second, t.map, p, e, x, cw, *omap);
  else if (e.res1 != nullptr || e.res2 != nullptr) le = cudaLaunchKernelEx(&cfg, conv_tc_kernel<EPI_GENERIC, true>, it->second, t.map, p, e, x, cw, *omap);
  else le = cudaLaunchKernelEx(&cfg, conv_tc_kernel<EPI_GENERIC, false>, it->second, t.map, p, e, x, cw, *omap);
  if (le != cudaSuccess) { tma.last_error = cudaGetErrorString(le); return -6; }
  (*launches)++;
  if (pc) { pc->resident += resident && !t.b_resident; pc->pinned += pinned; pc->staged += p.st_chunk != 0; }
  return 0;
}

}  // namespace bsr
