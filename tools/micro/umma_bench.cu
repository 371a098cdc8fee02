// Microbenchmark: cycles per tcgen05.mma (kind::f16, M=128, K=16, SS operands, 128B swizzle) as a function of N
// and of how many are issued back to back before one commit.  Operands are whatever is in shared memory.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../blindshadowremoval_b200/csrc/tc_common.cuh"
using namespace bsr;

__global__ void __launch_bounds__(128, 1) k(int n, int per_commit, int iters, long long* out, int mode) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + 16384, bar = base + 16384 + 32768, slot = bar + 8;
  volatile uint32_t* slot_p = reinterpret_cast<volatile uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(slot, 512);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot_p;
  if (threadIdx.x < 32) {
    const bool leader = elect_one();
    const uint32_t idesc = umma_idesc_bf16(128, n);
    const uint32_t a_lo = umma_desc_lo(sA), b_lo = umma_desc_lo(sB);
    uint32_t ph = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (leader) {
        for (int j = 0; j < per_commit; ++j) {
          const uint32_t koff = (uint32_t)(j & 3) * 2;
          if (mode == 0) umma_bf16_lo(tmem, a_lo + koff, b_lo + koff, idesc, 1u);
          else umma_bf16_lo(tmem + (uint32_t)((j & 1) * 256), a_lo + koff, b_lo + koff, idesc, 1u);   // alternate accumulators
        }
        umma_commit(bar);
      }
      __syncwarp();
      while (!mbar_try_wait(bar, ph)) {}
      ph ^= 1;
    }
    long long t1 = clock64();
    if (leader && blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int ns[] = {16, 32, 48, 64, 96, 128, 192, 256};
  for (int grid : {1, 148}) {
    for (int mode : {0, 1}) {
      for (int per : {4, 16, 64}) {
        printf("grid %3d mode %d per_commit %2d :", grid, mode, per);
        for (int n : ns) {
          const int iters = 200;
          k<<<grid, 128, 60 * 1024>>>(n, per, iters, d, mode);
          long long c = 0;
          cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
          printf("  N=%d:%6.1f", n, (double)c / (iters * per));
        }
        cudaError_t e = cudaDeviceSynchronize();
        printf("  [%s]\n", cudaGetErrorString(e));
      }
    }
  }
  return 0;
}
