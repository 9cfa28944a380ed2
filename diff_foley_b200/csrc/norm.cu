// norm.cu -- GroupNorm(32)+SiLU and LayerNorm producers of the fp16 GEMM operands.
//
// Both read the fp32 residual stream (channels-last) and emit the fp16 tile the following tensor-core
// GEMM / conv pulls in through TMA, so the normalised activation is written exactly once, at half
// width.  Statistics are two-pass (mean, then centred second moment) in fp32 like ATen's.
//
// Replaces: GroupNorm32 + SiLU (reference util.py:214-216, openai_unetmodel.py:201-203,225-228,
// 682-684), Normalize (attention_openai.py:76-77, eps 1e-6), nn.LayerNorm
// (attention_openai.py:203-205) and the th.cat of the skip connection (openai_unetmodel.py:736):
// the kernel normalises across the *concatenation* of two sources without materialising it in fp32.
#include "dfb_internal.h"
#include "dfb_ptx.cuh"

namespace dfb {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int THREADS>
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();  // protect `red` reuse across calls
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = (lane < THREADS / 32) ? red[lane] : 0.f;
  t = warp_sum(t);
  return t;  // every thread holds the total
}

constexpr int GN_THREADS = 512;

// grid = (32 groups, B).  The group's [HW x cpg] slab is staged in shared memory once.
__global__ void __launch_bounds__(GN_THREADS)
groupnorm_kernel(const float* __restrict__ src0, int C0, const float* __restrict__ src1, int C1,
                 int HW, const float* __restrict__ gamma, const float* __restrict__ beta,
                 float eps, int silu, __half* __restrict__ out, __half* __restrict__ raw_out) {
  extern __shared__ float slab[];  // [HW][cpg]
  __shared__ float red[GN_THREADS / 32];
  pdl_wait();
  pdl_launch_dependents();
  const int C = C0 + C1;
  const int cpg = C / 32;
  const int hp = cpg >> 1;  // channel pairs per pixel
  const int g = blockIdx.x, b = blockIdx.y;
  const int cbase = g * cpg;
  const int npairs = HW * hp;

  float s = 0.f;
  for (int i = threadIdx.x; i < npairs; i += GN_THREADS) {
    const int px = i / hp;
    const int c = cbase + 2 * (i - px * hp);
    float2 v;
    if (c < C0)
      v = *reinterpret_cast<const float2*>(src0 + ((size_t)b * HW + px) * C0 + c);
    else
      v = *reinterpret_cast<const float2*>(src1 + ((size_t)b * HW + px) * C1 + (c - C0));
    reinterpret_cast<float2*>(slab)[i] = v;
    s += v.x + v.y;
  }
  const float inv_n = 1.f / (float)(HW * cpg);
  const float mean = block_sum<GN_THREADS>(s, red) * inv_n;
  float q = 0.f;
  for (int i = threadIdx.x; i < npairs; i += GN_THREADS) {
    const float2 v = reinterpret_cast<const float2*>(slab)[i];
    const float dx = v.x - mean, dy = v.y - mean;
    q += dx * dx + dy * dy;
  }
  const float var = block_sum<GN_THREADS>(q, red) * inv_n;
  const float rstd = rsqrtf(var + eps);
  for (int i = threadIdx.x; i < npairs; i += GN_THREADS) {
    const int px = i / hp;
    const int c = cbase + 2 * (i - px * hp);
    const float2 v = reinterpret_cast<const float2*>(slab)[i];
    float y0 = (v.x - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
    float y1 = (v.y - mean) * rstd * __ldg(gamma + c + 1) + __ldg(beta + c + 1);
    if (silu) {
      y0 = y0 / (1.f + expf(-y0));
      y1 = y1 / (1.f + expf(-y1));
    }
    const size_t o = ((size_t)b * HW + px) * C + c;
    *reinterpret_cast<__half2*>(out + o) = __floats2half2_rn(y0, y1);
    if (raw_out != nullptr) *reinterpret_cast<__half2*>(raw_out + o) = __floats2half2_rn(v.x, v.y);
  }
}

int norm_init() {
  DFB_CUDA_OK(cudaFuncSetAttribute(groupnorm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   200 * 1024));
  return 0;
}

int groupnorm_launch(const float* src0, int C0, const float* src1, int C1, int B, int HW,
                     const float* gamma, const float* beta, float eps, int silu, __half* out,
                     __half* raw_out, cudaStream_t stream) {
  if (src1 == nullptr) C1 = 0;
  const int C = C0 + C1;
  if (C % 64 != 0 || (C0 & 1)) {
    set_error("groupnorm: channels must be a multiple of 64 (32 groups of an even width)");
    return -1;
  }
  const size_t smem = (size_t)HW * (C / 32) * sizeof(float);
  if (smem > 200 * 1024) {
    set_error("groupnorm: group slab of " + std::to_string(smem) + " bytes exceeds shared memory");
    return -1;
  }
  note("groupnorm", 0.0, (double)B * HW * C * (4.0 + 2.0 + (raw_out ? 2.0 : 0.0)), B * HW, C, 0, 1, 32 * B);
  DFB_CUDA_OK(launch_pdl(groupnorm_kernel, dim3(dim3(32, B)), dim3(GN_THREADS), smem, stream, src0, C0, src1, C1, HW, gamma, beta,
                                                              eps, silu, out, raw_out));
  DFB_CUDA_OK(cudaGetLastError());
  return 0;
}

// One warp per row of fp32 [rows, C]; three L1-resident passes (mean, centred variance, write).
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ src, int rows, int C, const float* __restrict__ gamma,
                 const float* __restrict__ beta, float eps, __half* __restrict__ out) {
  pdl_wait();
  pdl_launch_dependents();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* x = reinterpret_cast<const float4*>(src + (size_t)row * C);
  const int n4 = C >> 2;
  float s = 0.f;
  for (int i = lane; i < n4; i += 32) {
    const float4 v = x[i];
    s += (v.x + v.y) + (v.z + v.w);
  }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
  for (int i = lane; i < n4; i += 32) {
    const float4 v = x[i];
    const float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
  __half* o = out + (size_t)row * C;
  for (int i = lane; i < n4; i += 32) {
    const float4 v = x[i];
    const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + i);
    const float4 bt = __ldg(reinterpret_cast<const float4*>(beta) + i);
    const __half2 h0 = __floats2half2_rn((v.x - mean) * rstd * gm.x + bt.x,
                                         (v.y - mean) * rstd * gm.y + bt.y);
    const __half2 h1 = __floats2half2_rn((v.z - mean) * rstd * gm.z + bt.z,
                                         (v.w - mean) * rstd * gm.w + bt.w);
    uint2 u;
    u.x = *reinterpret_cast<const uint32_t*>(&h0);
    u.y = *reinterpret_cast<const uint32_t*>(&h1);
    *reinterpret_cast<uint2*>(o + 4 * i) = u;
  }
}

int layernorm_launch(const float* src, int rows, int C, const float* gamma, const float* beta,
                     float eps, __half* out, cudaStream_t stream) {
  if (C % 4 != 0) {
    set_error("layernorm: C must be a multiple of 4");
    return -1;
  }
  note("layernorm", 0.0, (double)rows * C * 6.0, rows, C, 0, 1, (rows + 7) / 8);
  DFB_CUDA_OK(launch_pdl(layernorm_kernel, dim3((rows + 7) / 8), dim3(256), 0, stream, src, rows, C, gamma, beta, eps, out));
  DFB_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace dfb
