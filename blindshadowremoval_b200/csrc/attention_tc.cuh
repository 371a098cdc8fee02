// Fused non-local attention on tcgen05 (h16 in, fp32 logits / softmax / accumulate):
//   y = softmax(theta . phi^T) . g      S = 1024 tokens, d = 128, one head, logits NOT scaled
// (/root/reference/model.py:51-53).  The 1024x1024 logit matrix never leaves the SM.
//
// One CTA per (image, 128-query tile).  Two passes over the 8 key tiles of 128:
//   pass 1: S_j = Q.K_j^T in TMEM (double buffered) -> running row max in registers;
//   pass 2: S_j again -> p = exp(s - max) -> h16 P tile in swizzled smem -> O += P.V_j in TMEM.
// Recomputing QK^T costs 1.5x the attention MMAs but needs no accumulator rescaling; the unscaled
// logits of this network make an exact (not running) max the safest choice as well.
// Layouts: QK[n][1024][256] = theta | phi (h16);  VT[n][128][1024] = g transposed (so every UMMA
// operand is K-major);  O[n][1024][128].
// Warp roles (320 threads): warps 0-7 softmax / epilogue (4 TMEM lane quarters x 2 column halves), warp 8 TMA,
// warp 9 TMEM alloc + MMA issue (single-thread roles get the highest warp ids: the arbiter favours high ids).
#pragma once
#include <map>
#include <tuple>
#include <utility>

#include "common.cuh"
#include "tc_common.cuh"

namespace bsr {

constexpr int AT_S = 1024, AT_D = 128, AT_BQ = 128, AT_BK = 128;
constexpr int AT_NK = AT_S / AT_BK;                   // 8 key tiles
constexpr uint32_t AT_TILE = 128 * 128 * 2;           // 32 KB: [128 rows][128 h16] as two 16 KB k-blocks
constexpr size_t kAttnTcSmem = 1024 + 7 * (size_t)AT_TILE + 192 + 1024 + 64;   // Q, K x2, V x2, P x2, barriers, exchange, residual-ring barriers

__device__ __forceinline__ void add_h16x16_attn(const uint4& a, const uint4& b, float* v) {
  const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float2 t = unpack_h16x2(w[i]);
    v[2 * i] += t.x;
    v[2 * i + 1] += t.y;
  }
}

#ifdef BSR_ROLE_TIMERS
#define AT_CLK() clock64()
#else
#define AT_CLK() 0ll
#endif
constexpr int AT_THREADS = 320;      // warps 0-7 softmax / epilogue, warp 8 TMA, warp 9 TMEM alloc + MMA issue

// fuse_w = 1: the NonLocalBlock output conv w (1x1, 128 -> 257, BN folded) and the ResBottleneck tail run here too:
//   out = LeakyReLU(x_in + y + W_w . O + b)      (model.py:56-59, 105-113)
// O (h16) goes to shared memory as the A operand of one more GEMM (N = 144 + 128 accumulator columns reuse the S / O
// TMEM columns, W_w lands in the Q/K buffers once the last S MMA has retired) and never reaches HBM.
// fuse_w = 2: same, with the block-tail epilogue staged through shared memory: the residual tiles (y, x_in) arrive by
//   TMA into the retired V stages (ring of two 64-channel batches) and the result leaves by TMA bulk stores from the
//   retired P[1] buffer.  Per-thread-row global accesses cost 32 L1 wavefronts per warp instruction (54 of them per
//   thread made the direct epilogue LSU-bound: 13.7 k of a CTA's 60 k cycles); the staged one touches shared memory only.
__global__ void __launch_bounds__(AT_THREADS, 1) attention_tc_kernel(const __grid_constant__ CUtensorMap tmQK,
                                                                     const __grid_constant__ CUtensorMap tmVT,
                                                                     const __grid_constant__ CUtensorMap tmW,
                                                                     const __grid_constant__ CUtensorMap tmR1,
                                                                     const __grid_constant__ CUtensorMap tmR2,
                                                                     const __grid_constant__ CUtensorMap tmOut,
                                                                     h16* __restrict__ o, const EpiParams e,
                                                                     const int fuse_w, int* errflag,
                                                                     long long* timers) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = base, sK = sQ + AT_TILE, sV = sK + 2 * AT_TILE, sP = sV + 2 * AT_TILE;
  const uint32_t bars = sP + 2 * AT_TILE;       // P is double-buffered: softmax(j+1) overlaps the PV MMA of tile j
  const uint32_t b_q = bars, b_kfull = bars + 8, b_kempty = bars + 24, b_vfull = bars + 40, b_vempty = bars + 56,
                 b_sfull = bars + 72, b_sempty = bars + 88, b_pfull = bars + 104, b_pempty = bars + 120,
                 b_ofull = bars + 136, b_wfull = bars + 144, b_a2full = bars + 152, b_d2full = bars + 160,
                 tmem_slot = bars + 168;
  uint8_t* smem_al = smem_raw + (base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_al + (tmem_slot - base));
  uint8_t* sP_gen = smem_al + (sP - base);
  float* xch = reinterpret_cast<float*>(smem_al + (bars + 192 - base));      // [2][128] row max / row sum exchange
  const uint32_t b_rfull = bars + 192 + 1024, b_rempty = b_rfull + 16;       // residual ring of the staged epilogue
  const uint32_t b_k2full = b_rfull + 32, b_k2empty = b_rfull + 48;          // pass 1: the V stages serve as K stages 2, 3
  uint8_t* sV_gen = smem_al + (sV - base);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * AT_BQ, n = blockIdx.y;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmQK);
    prefetch_tmap(&tmVT);
    if (fuse_w) prefetch_tmap(&tmW);
    mbar_init(b_q, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(b_kfull + 8 * s, 1);
      mbar_init(b_kempty + 8 * s, 1);
      mbar_init(b_vfull + 8 * s, 1);
      mbar_init(b_vempty + 8 * s, 1);
      mbar_init(b_sfull + 8 * s, 1);
      mbar_init(b_sempty + 8 * s, 8);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(b_pfull + 8 * s, 8);
      mbar_init(b_pempty + 8 * s, 1);
    }
    mbar_init(b_ofull, 1);
    mbar_init(b_wfull, 1);
    mbar_init(b_a2full, 8);
    mbar_init(b_d2full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(b_rfull + 8 * s, 1);
      mbar_init(b_rempty + 8 * s, 8);
      mbar_init(b_k2full + 8 * s, 1);
      mbar_init(b_k2empty + 8 * s, 1);
    }
    if (fuse_w == 2) {
      prefetch_tmap(&tmR1);
      prefetch_tmap(&tmR2);
      prefetch_tmap(&tmOut);
    }
    fence_barrier_init();
  }
  if (warp == 9) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  const uint32_t tS0 = tmem, tO = tmem + 256;
  // programmatic dependent launch: the prologue above overlapped the tail of the qkv conv; everything below
  // reads its output
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp == 8) {
    // ================= TMA producer (converged warp, elected lane issues) =================
    const bool leader = elect_one();
    if (leader) {
      mbar_expect_tx(b_q, AT_TILE);
      tma_load_3d(sQ, &tmQK, b_q, 0, q0, n);
      tma_load_3d(sQ + AT_TILE / 2, &tmQK, b_q, 64, q0, n);
    }
    bool ok = true;
    for (int it = 0; it < 2 * AT_NK && ok; ++it) {
      const int j = it & (AT_NK - 1);
      // Pass 1 (it < 8) has no V / P traffic, so the two V stages serve as K stages 2 and 3: four key tiles in flight
      // instead of two (with two, every S MMA waited a full L2 round trip for its tile).  Each stage is used twice in
      // pass 1, i.e. an even number of times, so the pass-2 parities of the K barriers are what they always were.
      const bool p1 = it < AT_NK;
      const int s = p1 ? (it & 3) : (it & 1), f = p1 ? (it >> 2) : (it >> 1);
      const uint32_t kfull = s < 2 ? b_kfull + 8 * s : b_k2full + 8 * (s - 2);
      const uint32_t kempty = s < 2 ? b_kempty + 8 * s : b_k2empty + 8 * (s - 2);
      const uint32_t kdst = s < 2 ? sK + s * AT_TILE : sV + (s - 2) * AT_TILE;
      ok = mbar_wait(kempty, (uint32_t)(f & 1) ^ 1u, errflag, 11);
      if (!ok) break;
      if (leader) {
        mbar_expect_tx(kfull, AT_TILE);
        tma_load_3d(kdst, &tmQK, kfull, 128, j * AT_BK, n);
        tma_load_3d(kdst + AT_TILE / 2, &tmQK, kfull, 192, j * AT_BK, n);
      }
      if (it >= AT_NK) {
        const int vs = j & 1, vf = j >> 1;
        // first V tile of a stage: the pass-1 S MMAs that read this stage as a K tile (second use) have retired
        if (vf == 0) ok = mbar_wait(b_k2empty + 8 * vs, 1u, errflag, 30);
        if (ok) ok = mbar_wait(b_vempty + 8 * vs, (uint32_t)(vf & 1) ^ 1u, errflag, 12);
        if (!ok) break;
        if (leader) {
          mbar_expect_tx(b_vfull + 8 * vs, AT_TILE);
          tma_load_3d(sV + vs * AT_TILE, &tmVT, b_vfull + 8 * vs, j * AT_BK, 0, n);
          tma_load_3d(sV + vs * AT_TILE + AT_TILE / 2, &tmVT, b_vfull + 8 * vs, j * AT_BK + 64, 0, n);
        }
      }
      __syncwarp();
    }
    if (fuse_w && ok) {
      // W_w (288 rows x 128 K, h16, 72 KB) replaces Q and the K stages once S_14 / S_15 have retired
      ok = mbar_wait(b_kempty, (uint32_t)((AT_NK) & 1) ^ 1u, errflag, 22);
      if (ok) ok = mbar_wait(b_kempty + 8, (uint32_t)((AT_NK) & 1) ^ 1u, errflag, 23);
      if (ok && leader) {
        mbar_expect_tx(b_wfull, 2 * 288 * 128);
        for (int kb = 0; kb < 2; ++kb) {
          tma_load_2d(sQ + kb * (288 * 128), &tmW, b_wfull, kb * 64, 0);
          tma_load_2d(sQ + kb * (288 * 128) + 144 * 128, &tmW, b_wfull, kb * 64, 144);
        }
      }
      __syncwarp();
    }
    if (fuse_w == 2 && ok) {
      // residual ring: batch b = channels 64b .. 64b+63 of y (res1) and x_in (res2) for this CTA's 128 pixels, into V
      // stage b & 1 (16 KB each); the stages are free once every PV MMA has retired
      ok = mbar_wait(b_ofull, 0, errflag, 27);
      const int pix0 = n * AT_S + q0;
      for (int b = 0; b < 4 && ok; ++b) {
        const int s = b & 1;
        if (b >= 2) ok = mbar_wait(b_rempty + 8 * s, 0, errflag, 28);
        if (!ok) break;
        if (leader) {
          mbar_expect_tx(b_rfull + 8 * s, AT_TILE);
          tma_load_2d(sV + s * AT_TILE, &tmR1, b_rfull + 8 * s, 64 * b, pix0);
          tma_load_2d(sV + s * AT_TILE + AT_TILE / 2, &tmR2, b_rfull + 8 * s, 64 * b, pix0);
        }
        __syncwarp();
      }
    }
  } else if (warp == 9) {
    // ================= MMA issuer (converged warp, elected lane issues) =================
    const bool leader = elect_one();
    const uint32_t idesc = umma_idesc_h16(128, 128);
    long long t_k = 0, t_se = 0, t_p = 0, t_v = 0, t_w = 0;
    const long long t0m = AT_CLK();
    bool ok = mbar_wait(b_q, 0, errflag, 13, true);
    auto issue_pv = [&](int j) -> bool {
      const int vs = j & 1, vf = j >> 1;
      const int ps = j & 1, pf = j >> 1;
      const long long c0 = AT_CLK();
      if (!mbar_wait(b_pfull + 8 * ps, (uint32_t)(pf & 1), errflag, 14, true)) return false;
      const long long c1 = AT_CLK();
      if (!mbar_wait(b_vfull + 8 * vs, (uint32_t)(vf & 1), errflag, 15, true)) return false;
      t_p += c1 - c0; t_v += AT_CLK() - c1;
      tc_fence_after();
      if (leader) {
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          const uint32_t a_lo = umma_desc_lo(sP + ps * AT_TILE + kb * (AT_TILE / 2));
          const uint32_t b_lo = umma_desc_lo(sV + vs * AT_TILE + kb * (AT_TILE / 2));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_h16_lo(tO, a_lo + 2 * k, b_lo + 2 * k, idesc, (j | kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(b_pempty + 8 * ps);
        umma_commit(b_vempty + 8 * vs);
      }
      __syncwarp();
      return true;
    };
    for (int it = 0; it < 2 * AT_NK && ok; ++it) {
      const int s = it & 1, f = it >> 1;                     // S accumulator stage (TMEM double buffer)
      const bool p1 = it < AT_NK;                            // pass 1: four K stages (see the producer)
      const int ks = p1 ? (it & 3) : s, kf = p1 ? (it >> 2) : f;
      const uint32_t kfull = ks < 2 ? b_kfull + 8 * ks : b_k2full + 8 * (ks - 2);
      const uint32_t kempty = ks < 2 ? b_kempty + 8 * ks : b_k2empty + 8 * (ks - 2);
      const uint32_t ksrc = ks < 2 ? sK + ks * AT_TILE : sV + (ks - 2) * AT_TILE;
      const long long c0 = AT_CLK();
      ok = mbar_wait(kfull, (uint32_t)(kf & 1), errflag, 16, true);
      if (!ok) break;
      const long long c1 = AT_CLK();
      ok = mbar_wait(b_sempty + 8 * s, (uint32_t)(f & 1) ^ 1u, errflag, 17, true);
      if (!ok) break;
      t_k += c1 - c0; t_se += AT_CLK() - c1;
      tc_fence_after();
      if (leader) {
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          const uint32_t a_lo = umma_desc_lo(sQ + kb * (AT_TILE / 2));
          const uint32_t b_lo = umma_desc_lo(ksrc + kb * (AT_TILE / 2));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_h16_lo(tS0 + (uint32_t)s * 128, a_lo + 2 * k, b_lo + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(kempty);
        umma_commit(b_sfull + 8 * s);
      }
      __syncwarp();
      if (it > AT_NK) ok = issue_pv(it - AT_NK - 1);
    }
    if (ok) ok = issue_pv(AT_NK - 1);
    if (leader) umma_commit(b_ofull);
    __syncwarp();
    const long long t1m = AT_CLK();
    if (fuse_w && ok) {
      ok = mbar_wait(b_wfull, 0, errflag, 24, true);
      if (ok) ok = mbar_wait(b_a2full, 0, errflag, 25, true);      // O (h16) is in sP[0]; S / O TMEM columns are drained
      t_w = AT_CLK() - t1m;
      tc_fence_after();
      if (ok && leader) {
        const uint32_t id144 = umma_idesc_h16(128, 144), id128 = umma_idesc_h16(128, 128);
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          const uint32_t a_lo = umma_desc_lo(sP + kb * (AT_TILE / 2));
          const uint32_t b_lo = umma_desc_lo(sQ + kb * (288 * 128));
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_h16_lo(tmem, a_lo + 2 * k, b_lo + 2 * k, id144, (kb | k) != 0 ? 1u : 0u);
            umma_h16_lo(tmem + 144, a_lo + 2 * k, b_lo + (144 * 8) + 2 * k, id128, (kb | k) != 0 ? 1u : 0u);
          }
        }
        umma_commit(b_d2full);
      }
      __syncwarp();
    }
#ifdef BSR_ROLE_TIMERS
    if (leader && blockIdx.x == 0 && blockIdx.y == 0 && timers) {
      timers[0] = AT_CLK() - t0m; timers[1] = t_k; timers[2] = t_se; timers[3] = t_p; timers[4] = t_v; timers[5] = t_w;
      timers[6] = t1m - t0m;
    }
#endif
  } else {
    // ================= softmax / epilogue: 8 warps = 4 TMEM lane quarters x 2 column halves =================
    const int q = warp & 3, h = warp >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const float kLog2e = 1.4426950408889634f;
    float mx = -INFINITY;
    bool ok = true;
    long long u_s1 = 0, u_s2 = 0, u_pe = 0;
    const long long t0s = AT_CLK();
    if (fuse_w) {
      // the residual rows (x_in, y) of this thread are needed ~15 us from now: pull them from HBM into L2 already
      const size_t pixp = (size_t)n * AT_S + q0 + row;
      const char* r1p = (const char*)((const h16*)e.res1 + pixp * e.res1_ld + h * 144);
      const char* r2p = (const char*)((const h16*)e.res2 + pixp * e.res2_ld + h * 144);
#pragma unroll
      for (int b = 0; b < 288; b += 128) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(r1p + b));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(r2p + b));
      }
    }
    // ---- pass 1: exact row max (each warp over its 64 columns, combined through smem at the end)
    for (int it = 0; it < AT_NK && ok; ++it) {
      const int s = it & 1, f = it >> 1;
      const long long c0 = AT_CLK();
      ok = mbar_wait(b_sfull + 8 * s, (uint32_t)(f & 1), errflag, 18);
      u_s1 += AT_CLK() - c0;
      if (!ok) break;
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 64; c += 32) {
        float v[32];
        tmem_ld32(tS0 + lane_addr + (uint32_t)(s * 128 + h * 64 + c), v);
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, v[i]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(b_sempty + 8 * s);
    }
    xch[h * 128 + row] = mx;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    mx = fmaxf(xch[row], xch[128 + row]);
    asm volatile("bar.sync 1, 256;" ::: "memory");
    // ---- pass 2: p = exp(s - max), P -> smem (h16, 128B-swizzled K-major), row sum
    const float mneg = -mx * kLog2e;
    float sum = 0.f;
    for (int j = 0; j < AT_NK && ok; ++j) {
      const int it = AT_NK + j, s = it & 1, f = it >> 1;
      const long long c0 = AT_CLK();
      ok = mbar_wait(b_sfull + 8 * s, (uint32_t)(f & 1), errflag, 19);
      if (!ok) break;
      const long long c1 = AT_CLK();
      const int ps = j & 1, pf = j >> 1;
      ok = mbar_wait(b_pempty + 8 * ps, (uint32_t)(pf & 1) ^ 1u, errflag, 20);
      if (!ok) break;
      u_s2 += c1 - c0; u_pe += AT_CLK() - c1;
      tc_fence_after();
      uint8_t* blk = sP_gen + ps * AT_TILE + h * (AT_TILE / 2) + row * 128;        // keys h*64 .. +63 = k-block h
#pragma unroll
      for (int c = 0; c < 64; c += 32) {
        float v[32];
        tmem_ld32(tS0 + lane_addr + (uint32_t)(s * 128 + h * 64 + c), v);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          v[i] = exp2f(fmaf(v[i], kLog2e, mneg));
          sum += v[i];
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int ch = (c >> 3) + g;                                // 16-byte chunk inside the 128-byte row
          uint4 w;
          w.x = pack_h16x2(v[8 * g + 0], v[8 * g + 1]);
          w.y = pack_h16x2(v[8 * g + 2], v[8 * g + 3]);
          w.z = pack_h16x2(v[8 * g + 4], v[8 * g + 5]);
          w.w = pack_h16x2(v[8 * g + 6], v[8 * g + 7]);
          *reinterpret_cast<uint4*>(blk + ((ch ^ (row & 7)) << 4)) = w;
        }
      }
      tc_fence_before();
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(b_sempty + 8 * s);
        mbar_arrive(b_pfull + 8 * ps);
      }
    }
    xch[h * 128 + row] = sum;
    // the K stage buffers are retired (S_15 was consumed above): stage the output conv's bias in shared memory
    float* bias_s = reinterpret_cast<float*>(smem_al + (sK + AT_TILE + 16384 - base));
    if (fuse_w)
      for (int i = threadIdx.x; i < 288; i += 256) bias_s[i] = __ldg(e.bias + i);
    asm volatile("bar.sync 1, 256;" ::: "memory");
    sum = xch[row] + xch[128 + row];
    // ---- epilogue: O / sum -> h16 (each warp its 64 output channels)
    const long long t1s = AT_CLK();
    if (ok) ok = mbar_wait(b_ofull, 0, errflag, 21);
    const long long t2s = AT_CLK();
    tc_fence_after();
    if (ok && !fuse_w) {
      const float inv = 1.f / sum;
      h16* dst = o + ((size_t)n * AT_S + q0 + row) * AT_D + h * 64;
#pragma unroll
      for (int c = 0; c < 64; c += 32) {
        float v[32];
        tmem_ld32(tO + lane_addr + (uint32_t)(h * 64 + c), v);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 w;
          w.x = pack_h16x2(v[8 * g + 0] * inv, v[8 * g + 1] * inv);
          w.y = pack_h16x2(v[8 * g + 2] * inv, v[8 * g + 3] * inv);
          w.z = pack_h16x2(v[8 * g + 4] * inv, v[8 * g + 5] * inv);
          w.w = pack_h16x2(v[8 * g + 6] * inv, v[8 * g + 7] * inv);
          *reinterpret_cast<uint4*>(dst + c + 8 * g) = w;
        }
      }
    }
    if (ok && fuse_w) {
      // O / sum -> h16 A operand in sP[0] (same swizzled K-major layout as P: this warp's 64 channels = k-block h)
      const float inv = 1.f / sum;
      uint8_t* blk = sP_gen + h * (AT_TILE / 2) + row * 128;
#pragma unroll
      for (int c = 0; c < 64; c += 32) {
        float v[32];
        tmem_ld32(tO + lane_addr + (uint32_t)(h * 64 + c), v);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int ch = (c >> 3) + g;
          uint4 w;
          w.x = pack_h16x2(v[8 * g + 0] * inv, v[8 * g + 1] * inv);
          w.y = pack_h16x2(v[8 * g + 2] * inv, v[8 * g + 3] * inv);
          w.z = pack_h16x2(v[8 * g + 4] * inv, v[8 * g + 5] * inv);
          w.w = pack_h16x2(v[8 * g + 6] * inv, v[8 * g + 7] * inv);
          *reinterpret_cast<uint4*>(blk + ((ch ^ (row & 7)) << 4)) = w;
        }
      }
      tc_fence_before();
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(b_a2full);
      if (fuse_w == 2) {
        // ---- staged block-tail epilogue (see kernel comment): batches of 64 channels; warp (q, h) owns rows 32q..32q+31 and
        // channels 32h..32h+31 of every batch; the last 16 (padding) channels 256..271 go the direct way (h == 0 warps)
        const size_t pix = (size_t)n * AT_S + q0 + row;
        uint4 ra[2], rb[2];
        if (h == 0) {
          const uint4* s1 = reinterpret_cast<const uint4*>((const h16*)e.res1 + pix * e.res1_ld + 256);
          const uint4* s2 = reinterpret_cast<const uint4*>((const h16*)e.res2 + pix * e.res2_ld + 256);
          ra[0] = s1[0]; ra[1] = s1[1]; rb[0] = s2[0]; rb[1] = s2[1];
        }
        ok = mbar_wait(b_d2full, 0, errflag, 26);
        tc_fence_after();
        const long long t4s = AT_CLK();
        auto finish16 = [&](float* v, const int c, const uint4& a0, const uint4& a1, const uint4& b0, const uint4& b1,
                            uint4& o0, uint4& o1) {
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c + i);
            v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
          }
          add_h16x16_attn(a0, a1, v);
          add_h16x16_attn(b0, b1, v);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], kLeaky * v[i]);
          o0.x = pack_h16x2(v[0], v[1]); o0.y = pack_h16x2(v[2], v[3]);
          o0.z = pack_h16x2(v[4], v[5]); o0.w = pack_h16x2(v[6], v[7]);
          o1.x = pack_h16x2(v[8], v[9]); o1.y = pack_h16x2(v[10], v[11]);
          o1.z = pack_h16x2(v[12], v[13]); o1.w = pack_h16x2(v[14], v[15]);
        };
        const uint32_t rowo = (uint32_t)row * 128u, sw = (uint32_t)row & 7u;
        for (int b = 0; b < 4 && ok; ++b) {
          const int s = b & 1;
          ok = mbar_wait(b_rfull + 8 * s, (uint32_t)(b >> 1), errflag, 29);
          if (!ok) break;
          const uint8_t* r1s = sV_gen + s * AT_TILE + rowo;
          const uint8_t* r2s = r1s + AT_TILE / 2;
          uint8_t* outs = sP_gen + AT_TILE + s * (AT_TILE / 2) + rowo;
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const int c = 64 * b + 32 * h + 16 * k;
            const uint32_t p0 = (((uint32_t)(4 * h + 2 * k)) ^ sw) << 4, p1 = (((uint32_t)(4 * h + 2 * k + 1)) ^ sw) << 4;
            float v[16];
            tmem_ld16(tmem + lane_addr + (uint32_t)c, v);
            const uint4 a0 = *reinterpret_cast<const uint4*>(r1s + p0), a1 = *reinterpret_cast<const uint4*>(r1s + p1);
            const uint4 b0 = *reinterpret_cast<const uint4*>(r2s + p0), b1 = *reinterpret_cast<const uint4*>(r2s + p1);
            uint4 o0, o1;
            finish16(v, c, a0, a1, b0, b1, o0, o1);
            *reinterpret_cast<uint4*>(outs + p0) = o0;
            *reinterpret_cast<uint4*>(outs + p1) = o1;
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(b_rempty + 8 * s);          // this warp is done with residual stage s
          // the bulk store issued one batch ago has finished reading the other staging half before anyone passes the
          // barrier and overwrites it in the next batch
          if (threadIdx.x == 0) bulk_wait_read0();
          asm volatile("bar.sync 1, 256;" ::: "memory");
          if (threadIdx.x == 0) {
            tma_store_2d(&tmOut, sP + AT_TILE + s * (AT_TILE / 2), 64 * b, (int)((size_t)n * AT_S + q0));
            bulk_commit();
          }
        }
        if (ok && h == 0) {
          float v[16];
          tmem_ld16(tmem + lane_addr + 256u, v);
          uint4 o0, o1;
          finish16(v, 256, ra[0], ra[1], rb[0], rb[1], o0, o1);
          uint4* d = reinterpret_cast<uint4*>((h16*)e.out + pix * e.out_ld + e.out_coff + 256);
          d[0] = o0;
          d[1] = o1;
        }
        if (threadIdx.x == 0) bulk_wait0();
#ifdef BSR_ROLE_TIMERS
        if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && timers) {
          timers[8] = t1s - t0s; timers[9] = u_s1; timers[10] = u_s2; timers[11] = u_pe; timers[12] = t2s - t1s;
          timers[13] = 0; timers[14] = t4s - t2s; timers[15] = AT_CLK() - t4s;
        }
#endif
      } else {
      // second GEMM's epilogue: warp half h owns accumulator columns [h*144, h*144 + (h ? 128 : 144))
      const size_t pix = (size_t)n * AT_S + q0 + row;
      const int c_begin = h * 144, c_end = h ? 272 : 144;
      const h16* r1 = (const h16*)e.res1 + pix * e.res1_ld;
      const h16* r2 = (const h16*)e.res2 + pix * e.res2_ld;
      h16* dst = (h16*)e.out + pix * e.out_ld + e.out_coff;
      // residual operands are fetched in batches of four 16-channel chunks (16 independent 16-byte loads in flight per
      // thread); the first batch is requested before the accumulator wait
      uint4 pa[4][2], pb[4][2];
      auto load_batch = [&](const int cb) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (cb + 16 * k < c_end) {
            const uint4* s1 = reinterpret_cast<const uint4*>(r1 + cb + 16 * k);
            const uint4* s2 = reinterpret_cast<const uint4*>(r2 + cb + 16 * k);
            pa[k][0] = s1[0]; pa[k][1] = s1[1];
            pb[k][0] = s2[0]; pb[k][1] = s2[1];
          }
        }
      };
      load_batch(c_begin);
      const long long t3s = AT_CLK();
      ok = mbar_wait(b_d2full, 0, errflag, 26);
      const long long t4s = AT_CLK();
#ifdef BSR_ROLE_TIMERS
      if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && timers) {
        timers[8] = t1s - t0s; timers[9] = u_s1; timers[10] = u_s2; timers[11] = u_pe; timers[12] = t2s - t1s;
        timers[13] = t3s - t2s; timers[14] = t4s - t3s;
      }
#endif
      tc_fence_after();
      if (ok) {
#pragma unroll 1
        for (int cb = c_begin; cb < c_end; cb += 64) {
          if (cb != c_begin) load_batch(cb);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int c = cb + 16 * k;
            if (c < c_end) {
              float v[16];
              tmem_ld16(tmem + lane_addr + (uint32_t)c, v);
#pragma unroll
              for (int i = 0; i < 16; i += 4) {
                const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c + i);
                v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
              }
              add_h16x16_attn(pa[k][0], pa[k][1], v);
              add_h16x16_attn(pb[k][0], pb[k][1], v);
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], kLeaky * v[i]);
              uint4 o0, o1;
              o0.x = pack_h16x2(v[0], v[1]); o0.y = pack_h16x2(v[2], v[3]);
              o0.z = pack_h16x2(v[4], v[5]); o0.w = pack_h16x2(v[6], v[7]);
              o1.x = pack_h16x2(v[8], v[9]); o1.y = pack_h16x2(v[10], v[11]);
              o1.z = pack_h16x2(v[12], v[13]); o1.w = pack_h16x2(v[14], v[15]);
              uint4* d = reinterpret_cast<uint4*>(dst + c);
              d[0] = o0;
              d[1] = o1;
            }
          }
        }
#ifdef BSR_ROLE_TIMERS
        if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && timers) timers[15] = AT_CLK() - t4s;
#endif
      }
      }      // fuse_w == 1 (direct epilogue)
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// w_map != nullptr: fused output conv (see kernel comment); `e` then carries bias / residuals / output of the block.
inline int launch_attention_tc(TmaEncoder& tma, const h16* qk, const h16* vt, h16* o, int n, int* errflag,
                               cudaStream_t st, const Knobs& kn, const CUtensorMap* w_map = nullptr,
                               const EpiParams* e = nullptr, int launch_index = 0) {
  static thread_local std::map<std::tuple<const void*, const void*, int>, std::pair<CUtensorMap, CUtensorMap>> cache;
  auto key = std::make_tuple((const void*)qk, (const void*)vt, n);
  auto it = cache.find(key);
  if (it == cache.end()) {
    CUtensorMap mq, mv;
    uint64_t dq[3] = {256, AT_S, (uint64_t)n}, sq[2] = {256 * 2, (uint64_t)AT_S * 256 * 2};
    uint32_t bq[3] = {64, 128, 1};
    if (!tma.encode_h16(&mq, (void*)qk, 3, dq, sq, bq, nullptr)) return -1;
    uint64_t dv[3] = {AT_S, AT_D, (uint64_t)n}, sv[2] = {(uint64_t)AT_S * 2, (uint64_t)AT_D * AT_S * 2};
    uint32_t bv[3] = {64, 128, 1};
    if (!tma.encode_h16(&mv, (void*)vt, 3, dv, sv, bv, nullptr)) return -2;
    if (cache.size() > 256) cache.clear();
    it = cache.emplace(key, std::make_pair(mq, mv)).first;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = dim3(AT_S / AT_BQ, n);
  cfg.blockDim = dim3(AT_THREADS);
  cfg.dynamicSmemBytes = kAttnTcSmem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = kn.no_pdl ? 0 : 1;
  EpiParams ep;
  memset(&ep, 0, sizeof ep);
  if (e) ep = *e;
  int fuse = (w_map != nullptr && e != nullptr) ? 1 : 0;
  // staged block-tail epilogue: [pixels x channels] maps of y (res1), x_in (res2) and the output, 64 x 128 boxes
  static thread_local std::map<std::tuple<const void*, int, int>, CUtensorMap> rcache;
  const CUtensorMap* rm[3] = {&it->second.first, &it->second.first, &it->second.first};
  if (fuse && !kn.no_tma_store && e->res1 && e->res2 && e->res1_ld >= 272 && e->res2_ld >= 272 &&
      e->out_ld >= e->out_coff + 272 && e->out_coff % 8 == 0 && e->out_ld % 8 == 0 && e->res1_ld % 8 == 0 && e->res2_ld % 8 == 0) {
    const void* bases[3] = {e->res1, e->res2, (const void*)((const h16*)e->out + e->out_coff)};
    const int lds[3] = {e->res1_ld, e->res2_ld, e->out_ld};
    bool okm = true;
    if (rcache.size() > 1024) rcache.clear();      // before this launch's lookups: pointers taken below stay valid
    for (int i = 0; i < 3 && okm; ++i) {
      auto rk = std::make_tuple(bases[i], lds[i], n);
      auto rt = rcache.find(rk);
      if (rt == rcache.end()) {
        CUtensorMap m;
        uint64_t d2[2] = {(uint64_t)lds[i], (uint64_t)n * AT_S}, s2[1] = {(uint64_t)lds[i] * 2};
        uint32_t b2[2] = {64, 128};
        if (!tma.encode_h16(&m, (void*)bases[i], 2, d2, s2, b2, nullptr)) { okm = false; break; }
        rt = rcache.emplace(rk, m).first;
      }
      rm[i] = &rt->second;
    }
    if (!okm) return -4;
    fuse = 2;
  }
  cudaError_t le = cudaLaunchKernelEx(&cfg, attention_tc_kernel, it->second.first, it->second.second,
                                      fuse ? *w_map : it->second.first, *rm[0], *rm[1], *rm[2], o, ep, fuse, errflag,
                                      reinterpret_cast<long long*>(errflag) + 16 + 16 * (launch_index & 63));
  if (le != cudaSuccess) { tma.last_error = cudaGetErrorString(le); return -3; }
  return 0;
}

inline int configure_tc_kernels_attn() {
  cudaError_t e = cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kAttnTcSmem);
  return e == cudaSuccess ? 0 : (int)e;
}

}  // namespace bsr
