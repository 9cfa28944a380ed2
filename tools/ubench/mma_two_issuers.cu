// Is the ~146-clock tcgen05.mma issue interval per CTA or per issuing warp?  W warps of one CTA each issue their own
// stream of 128 x N x 16 MMAs into their own accumulator; reports clocks per 64-wide k-block per warp and chip TFLOP/s.
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include "dfb_ptx.cuh"
using namespace dfb;

__global__ void __launch_bounds__(256) k(int N, int iters, int nissue, int same_operands, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[4];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (warp < nissue && lane == 0) {
    const uint32_t idesc = umma_idesc_f16(128, N);
    const uint32_t sa = smem_u32(smem + (same_operands ? 0 : warp * 49152));
    const uint64_t da = umma_desc_k_sw128(sa), db = umma_desc_k_sw128(sa + 16384);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) umma_f16_ss(tm + warp * N, da + 2 * kk, db + 2 * kk, idesc, 1u);
    }
    umma_commit(&bar[warp]);
    mbar_wait(&bar[warp], 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0 && warp == 0) out[0] = t1 - t0;
  }
  __syncwarp();
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 100000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int same = 0; same <= 1; ++same)
    for (int ni = 1; ni <= 4; ++ni)
      for (int N : {32, 64, 128}) {
        k<<<148, 256, 199 * 1024>>>(N, iters / 10, ni, same, d);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        k<<<148, 256, 199 * 1024>>>(N, iters, ni, same, d);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        long long c = 0; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
        printf("issuing warps %d (%s operands) N=%3d: %7.1f clk per k-block per warp, %6.1f ns wall -> %6.0f TFLOP/s chip %s\n", ni, same ? "same" : "own",
               N, (double)c / iters, ms * 1e6 / iters, 148.0 * ni * 2.0 * 128 * N * 64 * iters / (ms * 1e-3) / 1e12, e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
  return 0;
}
