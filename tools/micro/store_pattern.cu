// Micro-benchmark: how many L1 data-pipe wavefronts does a warp-wide global store cost, as a function of WHERE in
// their 128-byte lines the 32 lanes write?  (Motivation: the up3 / clr_up3 epilogue stores 32 B per lane at a 256 B lane
// stride, all lanes at the same offset inside their line.)  Run under ncu:
//   ncu --metrics l1tex__t_requests_pipe_lsu_mem_global_op_st.sum,l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_st.sum,\
//       l1tex__data_pipe_lsu_wavefronts_mem_lgds.sum,gpu__time_duration.sum  ./store_pattern
// Patterns (lane L of a warp, 128-byte output pixels):
//   0  32 B at pixel 2L, byte offset 32*cg          (the product epilogue: same in-line offset on every lane)
//   1  32 B at pixel 2L, byte offset 32*((L+j)%4)   (rotated: the four in-line offsets are spread over the lanes)
//   2  32 B at pixel L/4, byte offset 32*(L%4)      (a quad covers one line: 8 full lines per request)
//   3  16 B variants of 0 (two stores)               (the round-1 epilogue)
//   4  32 B at pixel L (stride 128 B), offset 32*cg
//   5  32 B at pixel L (stride 128 B), offset 32*((L+j)%4)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void st256(void* p, uint32_t v) {
  asm volatile("st.global.v8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st128(void* p, uint32_t v) {
  asm volatile("st.global.v4.b32 [%0], {%1, %1, %1, %1};" ::"l"(p), "r"(v) : "memory");
}

template <int PAT>
__global__ void __launch_bounds__(512) store_kernel(char* out, int tiles_per_cta, size_t tile_bytes) {
  const int warp = threadIdx.x >> 5, L = threadIdx.x & 31;
  const int q = warp & 3, cg = warp >> 2;
  for (int t = 0; t < tiles_per_cta; ++t) {
    // one tile = 128 input pixels -> 2 output rows of 256 pixels x 128 B (64 KB), as in up3
    char* base = out + ((size_t)blockIdx.x * tiles_per_cta + t) * tile_bytes;
    const int x = 32 * q + L;          // input pixel
#pragma unroll
    for (int j = 0; j < 4; ++j) {      // j = sub-pixel phase in patterns 0/1/3, chunk in 2
      const int py = j >> 1, px = j & 1;
      if (PAT == 0) st256(base + ((size_t)py * 256 + 2 * x + px) * 128 + 32 * cg, t);
      if (PAT == 1) {
        // lane L stores phase j of chunk (cg + L) % 4?  No: the warp owns all 4 chunks of phase (cg); j walks the chunks rotated
        const int ph_y = cg >> 1, ph_x = cg & 1;
        st256(base + ((size_t)ph_y * 256 + 2 * x + ph_x) * 128 + 32 * ((L + j) & 3), t);
      }
      if (PAT == 2) {
        const int ph_y = cg >> 1, ph_x = cg & 1;
        const int xx = 32 * q + 8 * j + (L >> 2);
        st256(base + ((size_t)ph_y * 256 + 2 * xx + ph_x) * 128 + 32 * (L & 3), t);
      }
      if (PAT == 3) {
        st128(base + ((size_t)py * 256 + 2 * x + px) * 128 + 32 * cg, t);
        st128(base + ((size_t)py * 256 + 2 * x + px) * 128 + 32 * cg + 16, t);
      }
      if (PAT == 4) st256(base + ((size_t)j * 128 + x) * 128 + 32 * cg, t);
      if (PAT == 5) st256(base + ((size_t)cg * 128 + x) * 128 + 32 * ((L + j) & 3), t);
    }
  }
}

int main() {
  const int ctas = 148, tiles = 222;
  const size_t tile_bytes = 65536, total = (size_t)ctas * tiles * tile_bytes;
  char* buf;
  if (cudaMalloc(&buf, total) != cudaSuccess) { printf("alloc failed\n"); return 1; }
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  for (int pat = 0; pat < 6; ++pat) {
    float best = 1e9f;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(a);
      switch (pat) {
        case 0: store_kernel<0><<<ctas, 512>>>(buf, tiles, tile_bytes); break;
        case 1: store_kernel<1><<<ctas, 512>>>(buf, tiles, tile_bytes); break;
        case 2: store_kernel<2><<<ctas, 512>>>(buf, tiles, tile_bytes); break;
        case 3: store_kernel<3><<<ctas, 512>>>(buf, tiles, tile_bytes); break;
        case 4: store_kernel<4><<<ctas, 512>>>(buf, tiles, tile_bytes); break;
        case 5: store_kernel<5><<<ctas, 512>>>(buf, tiles, tile_bytes); break;
      }
      cudaEventRecord(b);
      cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b);
      if (ms < best) best = ms;
    }
    printf("pattern %d: %.3f ms  %.0f GB/s  (%s)\n", pat, best, total / best / 1e6, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
