// attention.cu -- fused softmax(q k^T * scale) v for the UNet's self- and cross-attention.
//
// One CTA owns QB queries of one (sample, head); K/V stream through shared memory in blocks of 64
// keys with an online (running max / running sum) softmax, so the [B*8, Nq, Nk] score tensor the
// reference materialises in HBM (attention_openai.py:178-190: einsum -> softmax -> einsum, 32 MB per
// sample at the 16x64 level) never exists.  Inputs are the fp16 projections written by the QKV
// GEMM epilogue (head h at columns h*dpad of each row); the output is fp16 [B*Lq, heads*d], the A
// operand of the to_out GEMM -- the reference's '(b h) n d -> b n (h d)' rearrange is free.
//
// v0 (this file): fp32 SIMT math, warp-per-query, conflict-free padded K tile.  The tcgen05
// version (S and O accumulators in TMEM) replaces the inner products next; interface unchanged.
#include "dfb_internal.h"
#include "dfb_ptx.cuh"

namespace dfb {

constexpr int ATT_QB = 32;       // queries per CTA
constexpr int ATT_KB = 64;       // keys per shared-memory block
constexpr int ATT_THREADS = 256; // 8 warps, 4 queries each

__global__ void __launch_bounds__(ATT_THREADS)
attention_kernel(const __half* __restrict__ q, int ldq, const __half* __restrict__ k, int ldk,
                 const __half* __restrict__ v, int ldv, __half* __restrict__ out, int ldo, int heads,
                 int Lq, int Lk, int d, int dpad, float scale) {
  extern __shared__ __align__(16) uint8_t att_smem[];
  const int dk = d + 2;  // padded K row (odd number of 32-bit words -> conflict-free)
  __half* Ks = reinterpret_cast<__half*>(att_smem);           // [KB][d+2]
  __half* Vs = Ks + ATT_KB * dk;                              // [KB][d]
  __half* Qs = Vs + ATT_KB * d;                               // [QB][d]
  float* Ps = reinterpret_cast<float*>(Qs + ATT_QB * d);      // [8 warps][KB]

  const int bh = blockIdx.y;
  const int b = bh / heads, h = bh % heads;
  const int q0 = blockIdx.x * ATT_QB;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hd2 = d >> 1;

  // stage the query block
  for (int i = threadIdx.x; i < ATT_QB * hd2; i += ATT_THREADS) {
    const int r = i / hd2, c = i - r * hd2;
    __half2 val = __floats2half2_rn(0.f, 0.f);
    if (q0 + r < Lq)
      val = *reinterpret_cast<const __half2*>(q + ((size_t)b * Lq + q0 + r) * ldq + h * dpad + 2 * c);
    reinterpret_cast<__half2*>(Qs)[i] = val;
  }

  constexpr int QPW = ATT_QB / 8;  // queries per warp
  constexpr int MAXD32 = 5;        // d <= 160
  float m_run[QPW], l_run[QPW], acc[QPW][MAXD32];
#pragma unroll
  for (int i = 0; i < QPW; ++i) {
    m_run[i] = -INFINITY;
    l_run[i] = 0.f;
#pragma unroll
    for (int j = 0; j < MAXD32; ++j) acc[i][j] = 0.f;
  }

  for (int k0 = 0; k0 < Lk; k0 += ATT_KB) {
    __syncthreads();
    for (int i = threadIdx.x; i < ATT_KB * hd2; i += ATT_THREADS) {
      const int r = i / hd2, c = i - r * hd2;
      __half2 kv = __floats2half2_rn(0.f, 0.f), vv = kv;
      if (k0 + r < Lk) {
        const size_t row = (size_t)b * Lk + k0 + r;
        kv = *reinterpret_cast<const __half2*>(k + row * ldk + h * dpad + 2 * c);
        vv = *reinterpret_cast<const __half2*>(v + row * ldv + h * dpad + 2 * c);
      }
      *reinterpret_cast<__half2*>(Ks + r * dk + 2 * c) = kv;
      *reinterpret_cast<__half2*>(Vs + r * d + 2 * c) = vv;
    }
    __syncthreads();
#pragma unroll
    for (int qi = 0; qi < QPW; ++qi) {
      const int qr = warp * QPW + qi;
      const __half2* qrow = reinterpret_cast<const __half2*>(Qs + qr * d);
      // two keys per lane
      float s0 = 0.f, s1 = 0.f;
      const __half2* kr0 = reinterpret_cast<const __half2*>(Ks + lane * dk);
      const __half2* kr1 = reinterpret_cast<const __half2*>(Ks + (lane + 32) * dk);
      for (int c = 0; c < hd2; ++c) {
        const float2 qq = __half22float2(qrow[c]);
        const float2 a = __half22float2(kr0[c]);
        const float2 bb = __half22float2(kr1[c]);
        s0 = fmaf(qq.x, a.x, s0); s0 = fmaf(qq.y, a.y, s0);
        s1 = fmaf(qq.x, bb.x, s1); s1 = fmaf(qq.y, bb.y, s1);
      }
      s0 = (k0 + lane < Lk) ? s0 * scale : -INFINITY;
      s1 = (k0 + lane + 32 < Lk) ? s1 * scale : -INFINITY;
      float mx = fmaxf(s0, s1);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      const float m_new = fmaxf(m_run[qi], mx);
      const float corr = __expf(m_run[qi] - m_new);
      const float p0 = __expf(s0 - m_new), p1 = __expf(s1 - m_new);
      float ps = p0 + p1;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, o);
      l_run[qi] = l_run[qi] * corr + ps;
      m_run[qi] = m_new;
      float* pw = Ps + warp * ATT_KB;
      __syncwarp();
      pw[lane] = p0;
      pw[lane + 32] = p1;
      __syncwarp();
#pragma unroll
      for (int j = 0; j < MAXD32; ++j) {
        const int dd = lane + 32 * j;
        if (dd < d) {
          float a = acc[qi][j] * corr;
          for (int kk = 0; kk < ATT_KB; ++kk) a = fmaf(pw[kk], __half2float(Vs[kk * d + dd]), a);
          acc[qi][j] = a;
        }
      }
    }
  }
#pragma unroll
  for (int qi = 0; qi < QPW; ++qi) {
    const int qr = q0 + warp * QPW + qi;
    if (qr >= Lq) continue;
    const float inv = 1.f / l_run[qi];
#pragma unroll
    for (int j = 0; j < MAXD32; ++j) {
      const int dd = lane + 32 * j;
      if (dd < d) out[((size_t)b * Lq + qr) * ldo + h * d + dd] = __float2half_rn(acc[qi][j] * inv);
    }
  }
}

int attention_init() {
  DFB_CUDA_OK(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   96 * 1024));
  return 0;
}

int attention_launch(const __half* q, int ldq, const __half* k, int ldk, const __half* v, int ldv,
                     __half* out, int ldo, int B, int heads, int Lq, int Lk, int d, int dpad,
                     float scale, cudaStream_t stream) {
  if (d > 160 || (d & 1) || dpad < d) {
    set_error("attention: head dim must be even and <= 160");
    return -1;
  }
  const size_t smem = (size_t)(ATT_KB * (d + 2) + ATT_KB * d + ATT_QB * d) * sizeof(__half) +
                      8 * ATT_KB * sizeof(float);
  dim3 grid((Lq + ATT_QB - 1) / ATT_QB, B * heads);
  note("attention", 4.0 * B * heads * (double)Lq * Lk * d,
       2.0 * B * heads * ((double)Lq * d * 2 + (double)Lk * d * 2), Lq, Lk, d, 1, grid.x * grid.y);
  attention_kernel<<<grid, ATT_THREADS, smem, stream>>>(q, ldq, k, ldk, v, ldv, out, ldo, heads, Lq,
                                                        Lk, d, dpad, scale);
  DFB_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace dfb
