// stream_probe.cu -- how fast can a GEMM-shaped grid stream cold fp16 weights out of HBM?
// Compares the operand fetch of igemm_tcgen05_kernel (2-D TMA boxes of 128 rows x 128 B out of a
// row-major [N,K] matrix: 128 scattered 128-byte segments per stage) against 1-D bulk copies of
// pre-tiled contiguous 16 KB chunks, for several ring depths and grid sizes.  No MMA: the consumer
// releases a stage as soon as it lands, so this is the fetch ceiling of each scheme.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/stream_probe tools/stream_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)), "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"((uint64_t)src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

constexpr int TILE_BYTES = 128 * 64 * 2;  // BN=128 rows x 64 fp16

// mode 0: 2-D TMA box from row-major W[N,K]; mode 1: 1-D bulk from tiled [n_tile][kb][16 KB];
// mode 2: 1-D bulk, two 8 KB halves from [n64][kb][8 KB]
__global__ void __launch_bounds__(64) probe(const __grid_constant__ CUtensorMap tm, const uint8_t* tiled,
                                            int kb_total, int splits, int stages, int mode) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + stages * TILE_BYTES);
  uint64_t* empty = full + stages;
  const int nt = blockIdx.x, z = blockIdx.y;
  const int kb0 = (int)((long)z * kb_total / splits), kb1 = (int)((long)(z + 1) * kb_total / splits);
  const int nkb = kb1 - kb0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 0; i < nkb; ++i) {
      const int s = i % stages;
      if (i >= stages) mbar_wait(&empty[s], ((i / stages) & 1) ^ 1);
      mbar_expect_tx(&full[s], TILE_BYTES);
      uint8_t* dst = smem + s * TILE_BYTES;
      const int kb = kb0 + i;
      if (mode == 0) {
        tma_load_2d(dst, &tm, &full[s], kb * 64, nt * 128);
      } else if (mode == 1) {
        bulk_load_1d(dst, tiled + ((size_t)nt * kb_total + kb) * TILE_BYTES, TILE_BYTES, &full[s]);
      } else {
        bulk_load_1d(dst, tiled + ((size_t)(2 * nt) * kb_total + kb) * 8192, 8192, &full[s]);
        bulk_load_1d(dst + 8192, tiled + ((size_t)(2 * nt + 1) * kb_total + kb) * 8192, 8192, &full[s]);
      }
    }
  } else if (threadIdx.x == 32) {
    for (int i = 0; i < nkb; ++i) {
      const int s = i % stages;
      mbar_wait(&full[s], (i / stages) & 1);
      mbar_arrive(&empty[s]);
    }
  }
}

// plain vectorised read with a full grid: the ceiling any scheme is compared with
__global__ void read_all(const uint4* p, size_t n, uint4* sink) {
  uint4 acc = make_uint4(0, 0, 0, 0);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const uint4 v = p[i];
    acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
  }
  if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x12345678u) sink[0] = acc;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int N = 1280, K = 11520, NBUF = 8;  // 29.5 MB per matrix, 236 MB rotating (> 126 MB L2)
  const size_t bytes = (size_t)N * K * 2;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  PFN_encodeTiled enc = (PFN_encodeTiled)p;
  std::vector<uint8_t*> bufs(NBUF);
  for (auto& b : bufs) { CK(cudaMalloc(&b, bytes)); CK(cudaMemset(b, 1, bytes)); }
  uint4* sink; CK(cudaMalloc(&sink, 64));
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int kb_total = K / 64, tiles_n = N / 128;

  {  // ceiling
    for (int it = 0; it < 2; ++it)
      for (int b = 0; b < NBUF; ++b) read_all<<<148 * 8, 256>>>((const uint4*)bufs[b], bytes / 16, sink);
    CK(cudaEventRecord(e0));
    for (int b = 0; b < NBUF; ++b) read_all<<<148 * 8, 256>>>((const uint4*)bufs[b], bytes / 16, sink);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("read_all (148x8 CTAs x 256 thr, LDG.128)            %8.2f us/matrix  %7.0f GB/s\n", ms * 1e3 / NBUF, bytes / (ms / NBUF * 1e-3) * 1e-9);
  }
  const CUtensorMapL2promotion promos[3] = {CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE};
  const char* promo_names[3] = {"256B", "128B", "none"};
  for (int mode = 0; mode < 3; ++mode) {
    for (int pi = 0; pi < (mode == 0 ? 3 : 1); ++pi) {
      std::vector<CUtensorMap> tms(NBUF);
      for (int b = 0; b < NBUF; ++b) {
        cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)N};
        cuuint64_t gstr[1] = {(cuuint64_t)K * 2};
        cuuint32_t box[2] = {64, 128}, es[2] = {1, 1};
        CUresult r = enc(&tms[b], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, bufs[b], gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, promos[pi], CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
      }
      for (int splits : {7, 14, 28}) {
        for (int stages : {3, 6, 12}) {
          const size_t smem = (size_t)stages * TILE_BYTES + 2 * stages * 8 + 1024;
          if (smem > 220 * 1024) continue;
          dim3 grid(tiles_n, splits);
          for (int it = 0; it < 2; ++it)
            for (int b = 0; b < NBUF; ++b) probe<<<grid, 64, smem>>>(tms[b], bufs[b], kb_total, splits, stages, mode);
          CK(cudaGetLastError());
          CK(cudaEventRecord(e0));
          for (int b = 0; b < NBUF; ++b) probe<<<grid, 64, smem>>>(tms[b], bufs[b], kb_total, splits, stages, mode);
          CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
          float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
          printf("mode %d (%s) promo %-4s ctas %3d stages %2d   %8.2f us/matrix  %7.0f GB/s\n", mode,
                 mode == 0 ? "2-D TMA 128x128B rows " : (mode == 1 ? "1-D bulk 16 KB tiles  " : "1-D bulk 2 x 8 KB     "),
                 mode == 0 ? promo_names[pi] : "-", tiles_n * splits, stages, ms * 1e3 / NBUF, bytes / (ms / NBUF * 1e-3) * 1e-9);
        }
      }
    }
  }
  return 0;
}
