// tcgen05.mma issue floor in SS mode (both operands in shared memory, 128-byte swizzled K-major tiles):
// cycles per 128 x N x 16 fp16 MMA for N = 32..256, one or two CTAs per SM, with and without concurrent
// shared-memory write traffic from the other warps (a stand-in for the TMA fill of the operand ring).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I diff_foley_b200/csrc -o gpurun_out/mma_floor tools/ubench/mma_floor.cu
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include "dfb_ptx.cuh"
using namespace dfb;

// whole-warp (uniform control flow) forms: one elected lane issues (elect_one() is dfb_ptx.cuh's)
__device__ __forceinline__ void MMA(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (elect_one()) umma_f16_ss(d, a, b, idesc, acc);
}
__device__ __forceinline__ void COMMIT(uint64_t* bar) {
  if (elect_one()) umma_commit(bar);
}

__global__ void __launch_bounds__(256) mma_floor(int N, int iters, int writers, int stage_bytes, int nstages, int commit_mode, int nacc, int tmem_cols, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint64_t ebar[8];
  __shared__ uint32_t slot;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); for (int i = 0; i < 8; ++i) mbar_init(&ebar[i], 1); fence_mbar_init(); stop = 0; }
  if (warp == 0) { tmem_alloc(&slot, tmem_cols); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
#ifdef ELECT
  if (warp == 0) {
    {
      const uint32_t idesc = umma_idesc_f16(128, N);
      const long long t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        const uint32_t sa = smem_u32(smem + (i % nstages) * stage_bytes);
        const uint64_t da = umma_desc_k_sw128(sa), db = umma_desc_k_sw128(sa + 16384);
#pragma unroll
        for (int k = 0; k < 4; ++k) MMA(tm + (uint32_t)(((i * 4 + k) & (nacc - 1)) * N), da + 2 * k, db + 2 * k, idesc, 1u);
        // commit_mode 1: one commit per k-block (8 rotating barriers, never waited on); 2: additionally wait for the
        // commit issued 6 k-blocks ago (what the TMA producer of the ring does before it refills that slot)
        if (commit_mode >= 1) COMMIT(&ebar[i & 7]);
        if (commit_mode == 2 && i >= 6) mbar_wait(&ebar[(i - 6) & 7], ((i - 6) >> 3) & 1);
      }
      COMMIT(&bar);
      mbar_wait(&bar, 0);
      const long long t1 = clock64();
      stop = 1;
      if (blockIdx.x == 0 && lane == 0) out[0] = t1 - t0;
    }
    __syncwarp();
  }
#else
  if (warp == 0) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_f16(128, N);
      const long long t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        const uint32_t sa = smem_u32(smem + (i % nstages) * stage_bytes);
        const uint64_t da = umma_desc_k_sw128(sa), db = umma_desc_k_sw128(sa + 16384);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_ss(tm + (uint32_t)(((i * 4 + k) & (nacc - 1)) * N), da + 2 * k, db + 2 * k, idesc, 1u);
        if (commit_mode >= 1) umma_commit(&ebar[i & 7]);
        if (commit_mode == 2 && i >= 6) mbar_wait(&ebar[(i - 6) & 7], ((i - 6) >> 3) & 1);
      }
      umma_commit(&bar);
      mbar_wait(&bar, 0);
      const long long t1 = clock64();
      stop = 1;
      if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    __syncwarp();
  }
#endif
  if (warp >= 1 && warp <= writers) {
    // 16-byte stores sweeping a separate 32 KB region, conflict-free
    const uint32_t base = smem_u32(smem + nstages * stage_bytes) + threadIdx.x * 16;
    int j = 0;
    while (!stop) {
      sts_f4(base + ((j & 7) << 12), 1.f, 2.f, 3.f, 4.f);
      ++j;
    }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, tmem_cols); }
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(mma_floor, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 100000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int ctas_per_sm = 1; ctas_per_sm <= 2; ++ctas_per_sm)
    for (int cm = 0; cm <= 2; cm += 2)
     for (int nacc = 1; nacc <= 4; nacc *= 2)
      for (int N : {32, 64, 128, 256}) {
        if (nacc * N * ctas_per_sm > 512) continue;
        const int writers = 0;
        const int stage = 16384 + N * 128, nst = 2;
        const int smem = nst * stage + 32768 + 1024 + (ctas_per_sm == 1 && N < 256 ? 80 * 1024 : 0);   // pad: force 1 CTA / SM
        mma_floor<<<148 * ctas_per_sm, 256, smem>>>(N, iters, writers, stage, nst, cm, nacc, 512 / ctas_per_sm, d);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        mma_floor<<<148 * ctas_per_sm, 256, smem>>>(N, iters, writers, stage, nst, cm, nacc, 512 / ctas_per_sm, d);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        long long c = 0; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
        printf("ctas/SM %d  commit mode %d  accumulators %d  N=%3d: %7.1f clk per MMA (128xNx16), %7.1f clk = %6.1f ns (wall) per 64-wide k-block per CTA -> %6.0f TFLOP/s chip  %s\n",
               ctas_per_sm, cm, nacc, N, (double)c / (4.0 * iters), (double)c / iters, ms * 1e6 / iters,
               148.0 * ctas_per_sm * 2.0 * 128 * N * 64 * iters / (ms * 1e-3) / 1e12, e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
  return 0;
}
