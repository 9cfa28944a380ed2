// backward.cu -- the kernels the double-guidance classifier's  d/dx log p(x_t, t, video)  needs besides the
// tensor-core GEMM (reference: ddim.py:333-341 calls torch.autograd.grad through Classifier_Backbone,
// alignment_backbone.py:417-686).  Backward-data of every conv / Linear is the same implicit-GEMM kernel with
// rotated / transposed packed weights (igemm_tcgen05.cu); what is left are the pointwise / reduction
// backward passes below.  Only d/dx is needed (no weight gradients).  Activations are channels-last fp32
// [B, HW, C]; gradients that feed a GEMM are emitted in fp16 (its A operand), residual-path sums in fp32.
//   groupnorm_bwd      GroupNorm(32)(+SiLU) backward, statistics recomputed from the saved input
//   layernorm_bwd      LayerNorm backward (warp per row)
//   attention_bwd_q/kv softmax(q k^T s) v backward per (sample, head): deterministic, no atomics --
//                      one kernel owns query rows (dq + log-sum-exp), one owns key rows (dk, dv)
//   geglu_fwd / _bwd   a * gelu_erf(g) and its derivative (attention_openai.py:37-44)
//   col2im_s2          gather form of the stride-2 3x3 conv's backward-data scatter
//   classifier_head    avg-pool + Linear(C -> 1) + sigmoid + log and the gradient seed (alignment_backbone.py:676-686)
#include <algorithm>

#include "dfb_internal.h"
#include "dfb_ptx.cuh"

namespace dfb {

__device__ __forceinline__ float warp_sum_b(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// block-wide sum of two values (blockDim.x <= 1024, multiple of 32); every thread gets the totals
__device__ __forceinline__ void block_sum2(float& a, float& b, float* red) {
  a = warp_sum_b(a);
  b = warp_sum_b(b);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) { red[warp] = a; red[32 + warp] = b; }
  __syncthreads();
  a = warp_sum_b(lane < nw ? red[lane] : 0.f);
  b = warp_sum_b(lane < nw ? red[32 + lane] : 0.f);
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + __expf(-x)); }

// ---- GroupNorm(32) [+ SiLU] backward.  One CTA per (group, sample); y = act(xhat * gamma + beta),
// xhat = (x - mean) * rstd over the group's HW x cpg elements.  g = dy * act'(pre) * gamma;
// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)).  add (optional fp32) is summed into dx (residual path).
__global__ void __launch_bounds__(256)
groupnorm_bwd_kernel(const float* __restrict__ x, int C, int HW, const float* __restrict__ gamma,
                     const float* __restrict__ beta, float eps, int silu, const float* __restrict__ dy,
                     const float* __restrict__ add, float* __restrict__ dx32, __half* __restrict__ dx16) {
  __shared__ float red[64];
  pdl_wait();
  pdl_launch_dependents();
  const int cpg = C >> 5, g = blockIdx.x, b = blockIdx.y;
  const int n = HW * cpg;
  const size_t base = (size_t)b * HW * C + (size_t)g * cpg;
  float s = 0.f, q = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += x[base + (size_t)(i / cpg) * C + (i % cpg)];
  float dummy = 0.f;
  block_sum2(s, dummy, red);
  const float mean = s / (float)n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float d = x[base + (size_t)(i / cpg) * C + (i % cpg)] - mean;
    q += d * d;
  }
  dummy = 0.f;
  block_sum2(q, dummy, red);
  const float rstd = rsqrtf(q / (float)n + eps);
  float s1 = 0.f, s2 = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int c = i % cpg;
    const size_t o = base + (size_t)(i / cpg) * C + c;
    const float xh = (x[o] - mean) * rstd;
    const float gm = gamma[g * cpg + c];
    float gg = dy[o] * gm;
    if (silu) {
      const float pre = fmaf(xh, gm, beta[g * cpg + c]);
      const float sg = sigmoid_f(pre);
      gg *= sg * (1.f + pre * (1.f - sg));
    }
    s1 += gg;
    s2 += gg * xh;
  }
  block_sum2(s1, s2, red);
  const float m1 = s1 / (float)n, m2 = s2 / (float)n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int c = i % cpg;
    const size_t o = base + (size_t)(i / cpg) * C + c;
    const float xh = (x[o] - mean) * rstd;
    const float gm = gamma[g * cpg + c];
    float gg = dy[o] * gm;
    if (silu) {
      const float pre = fmaf(xh, gm, beta[g * cpg + c]);
      const float sg = sigmoid_f(pre);
      gg *= sg * (1.f + pre * (1.f - sg));
    }
    float d = rstd * (gg - m1 - xh * m2);
    if (add != nullptr) d += add[o];
    if (dx32 != nullptr) dx32[o] = d;
    if (dx16 != nullptr) dx16[o] = __float2half_rn(d);
  }
}

int groupnorm_bwd_launch(const float* x, int C, int B, int HW, const float* gamma, const float* beta, float eps,
                         int silu, const float* dy, const float* add, float* dx32, __half* dx16, cudaStream_t stream) {
  if (C % 32) { set_error("groupnorm_bwd: channels must be a multiple of 32"); return -1; }
  note("groupnorm_bwd", 0.0, (double)B * HW * C * 24.0);
  DFB_CUDA_OK(launch_pdl(groupnorm_bwd_kernel, dim3(32, B), dim3(256), 0, stream, x, C, HW, gamma, beta, eps, silu, dy,
                         add, dx32, dx16));
  DFB_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---- LayerNorm backward, one warp per row of fp32 [rows, C]
__global__ void __launch_bounds__(128)
layernorm_bwd_kernel(const float* __restrict__ x, int rows, int C, const float* __restrict__ gamma, float eps,
                     const float* __restrict__ dy, const float* __restrict__ add, float* __restrict__ dx32,
                     __half* __restrict__ dx16) {
  pdl_wait();
  pdl_launch_dependents();
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + (size_t)row * C;
  const float* dr = dy + (size_t)row * C;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += xr[c];
  const float mean = warp_sum_b(s) / (float)C;
  float q = 0.f;
  for (int c = lane; c < C; c += 32) { const float d = xr[c] - mean; q += d * d; }
  const float rstd = rsqrtf(warp_sum_b(q) / (float)C + eps);
  float s1 = 0.f, s2 = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float gg = dr[c] * gamma[c];
    s1 += gg;
    s2 += gg * (xr[c] - mean) * rstd;
  }
  const float m1 = warp_sum_b(s1) / (float)C, m2 = warp_sum_b(s2) / (float)C;
  for (int c = lane; c < C; c += 32) {
    const float xh = (xr[c] - mean) * rstd;
    float d = rstd * (dr[c] * gamma[c] - m1 - xh * m2);
    const size_t o = (size_t)row * C + c;
    if (add != nullptr) d += add[o];
    if (dx32 != nullptr) dx32[o] = d;
    if (dx16 != nullptr) dx16[o] = __float2half_rn(d);
  }
}

int layernorm_bwd_launch(const float* x, int rows, int C, const float* gamma, float eps, const float* dy,
                         const float* add, float* dx32, __half* dx16, cudaStream_t stream) {
  note("layernorm_bwd", 0.0, (double)rows * C * 20.0);
  DFB_CUDA_OK(launch_pdl(layernorm_bwd_kernel, dim3((rows + 3) / 4), dim3(128), 0, stream, x, rows, C, gamma, eps, dy, add,
                         dx32, dx16));
  DFB_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---- attention backward.  q/k/v fp16 rows with head h at columns h*d (d <= 64, multiple of 8); o (forward
// output, fp16 [B*Lq, ldo]) and dO (fp32 [B*Lq, lddo]) with head h at columns h*d.
//   P = softmax(s * Q K^T);  D_i = sum_d dO_id O_id;  dS = P o (dO V^T - D) * s;  dQ = dS K;  dK = dS^T Q;  dV = P^T dO
// attention_bwd_q: CTA = 32 query rows of one (sample, head); two sweeps over the keys (row max / sum, then the
// gradient); writes dq (fp16) and the rows' log-sum-exp + D for attention_bwd_kv, whose CTAs own 32 key rows and
// sweep the queries.  No atomics: every output element has one writer, sums run in a fixed order.
constexpr int AB_T = 32;   // rows per CTA / keys per chunk
constexpr int AB_DMAX = 64;

__global__ void __launch_bounds__(256)
attention_bwd_q_kernel(const __half* __restrict__ q, int ldq, const __half* __restrict__ k, int ldk,
                       const __half* __restrict__ v, int ldv, const __half* __restrict__ o, int ldo,
                       const float* __restrict__ dO, int lddo, int Lq, int Lk, int d, float scale,
                       __half* __restrict__ dq, int lddq, float* __restrict__ lse, float* __restrict__ Dv) {
  __shared__ float sq[AB_T][AB_DMAX + 1], sdo[AB_T][AB_DMAX + 1], sk[AB_T][AB_DMAX + 1], sv[AB_T][AB_DMAX + 1];
  __shared__ float sds[8][4][AB_T];
  pdl_wait();
  pdl_launch_dependents();
  const int q0 = blockIdx.x * AB_T, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nh = gridDim.y;
  // load the CTA's query rows and dO rows (fp32 in shared memory)
  for (int i = threadIdx.x; i < AB_T * d; i += 256) {
    const int r = i / d, c = i % d;
    const int row = q0 + r;
    sq[r][c] = row < Lq ? __half2float(q[((size_t)b * Lq + row) * ldq + h * d + c]) : 0.f;
    sdo[r][c] = row < Lq ? dO[((size_t)b * Lq + row) * lddo + h * d + c] : 0.f;
  }
  __syncthreads();
  // D_r = sum_d dO o  (warp w owns rows 4w..4w+3)
  float Dr[4], mx[4], sm[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int r = warp * 4 + j, row = q0 + r;
    float t = 0.f;
    if (row < Lq)
      for (int c = lane; c < d; c += 32) t += sdo[r][c] * __half2float(o[((size_t)b * Lq + row) * ldo + h * d + c]);
    Dr[j] = warp_sum_b(t);
    mx[j] = -INFINITY;
    sm[j] = 0.f;
  }
  const int nchunk = (Lk + AB_T - 1) / AB_T;
  // ---- sweep 1: running max / sum of exp (lane = key within the chunk)
  for (int ch = 0; ch < nchunk; ++ch) {
    __syncthreads();
    for (int i = threadIdx.x; i < AB_T * d; i += 256) {
      const int r = i / d, c = i % d, key = ch * AB_T + r;
      sk[r][c] = key < Lk ? __half2float(k[((size_t)b * Lk + key) * ldk + h * d + c]) : 0.f;
    }
    __syncthreads();
    const bool kv = (ch * AB_T + lane) < Lk;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = warp * 4 + j;
      float s = 0.f;
      for (int c = 0; c < d; ++c) s = fmaf(sq[r][c], sk[lane][c], s);
      s = kv ? s * scale : -INFINITY;
      float m = s;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
      const float mn = fmaxf(mx[j], m);
      const float e = kv ? __expf(s - mn) : 0.f;
      sm[j] = sm[j] * __expf(mx[j] - mn) + warp_sum_b(e);
      mx[j] = mn;
    }
  }
  float ls[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    ls[j] = mx[j] + __logf(sm[j]);
    const int row = q0 + warp * 4 + j;
    if (lane == 0 && row < Lq) {
      lse[((size_t)b * nh + h) * Lq + row] = ls[j];
      Dv[((size_t)b * nh + h) * Lq + row] = Dr[j];
    }
  }
  // ---- sweep 2: dS and dq += dS K (lane = key for dS, lane = feature for the accumulation)
  float acc[4][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
  for (int ch = 0; ch < nchunk; ++ch) {
    __syncthreads();
    for (int i = threadIdx.x; i < AB_T * d; i += 256) {
      const int r = i / d, c = i % d, key = ch * AB_T + r;
      const bool ok = key < Lk;
      sk[r][c] = ok ? __half2float(k[((size_t)b * Lk + key) * ldk + h * d + c]) : 0.f;
      sv[r][c] = ok ? __half2float(v[((size_t)b * Lk + key) * ldv + h * d + c]) : 0.f;
    }
    __syncthreads();
    const bool kv = (ch * AB_T + lane) < Lk;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = warp * 4 + j;
      float s = 0.f, dp = 0.f;
      for (int c = 0; c < d; ++c) {
        s = fmaf(sq[r][c], sk[lane][c], s);
        dp = fmaf(sdo[r][c], sv[lane][c], dp);
      }
      const float p = kv ? __expf(s * scale - ls[j]) : 0.f;
      sds[warp][j][lane] = p * (dp - Dr[j]) * scale;
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int c = lane + 32 * u;
        if (c < d) {
          float a = acc[j][u];
          for (int key = 0; key < AB_T; ++key) a = fmaf(sds[warp][j][key], sk[key][c], a);
          acc[j][u] = a;
        }
      }
    }
    __syncwarp();
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int row = q0 + warp * 4 + j;
    if (row < Lq) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int c = lane + 32 * u;
        if (c < d) dq[((size_t)b * Lq + row) * lddq + h * d + c] = __float2half_rn(acc[j][u]);
      }
    }
  }
}

__global__ void __launch_bounds__(256)
attention_bwd_kv_kernel(const __half* __restrict__ q, int ldq, const __half* __restrict__ k, int ldk,
                        const __half* __restrict__ v, int ldv, const float* __restrict__ dO, int lddo, int Lq, int Lk,
                        int d, float scale, const float* __restrict__ lse, const float* __restrict__ Dv,
                        __half* __restrict__ dk, int lddk, __half* __restrict__ dv, int lddv) {
  __shared__ float sk[AB_T][AB_DMAX + 1], sv[AB_T][AB_DMAX + 1], sq[AB_T][AB_DMAX + 1], sdo[AB_T][AB_DMAX + 1];
  __shared__ float sp[8][4][AB_T], sds[8][4][AB_T], sl[AB_T], sD[AB_T];
  pdl_wait();
  pdl_launch_dependents();
  const int k0 = blockIdx.x * AB_T, h = blockIdx.y, b = blockIdx.z, nh = gridDim.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < AB_T * d; i += 256) {
    const int r = i / d, c = i % d, key = k0 + r;
    const bool ok = key < Lk;
    sk[r][c] = ok ? __half2float(k[((size_t)b * Lk + key) * ldk + h * d + c]) : 0.f;
    sv[r][c] = ok ? __half2float(v[((size_t)b * Lk + key) * ldv + h * d + c]) : 0.f;
  }
  float ak[4][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}}, av[4][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
  const int nchunk = (Lq + AB_T - 1) / AB_T;
  for (int ch = 0; ch < nchunk; ++ch) {
    __syncthreads();
    for (int i = threadIdx.x; i < AB_T * d; i += 256) {
      const int r = i / d, c = i % d, row = ch * AB_T + r;
      const bool ok = row < Lq;
      sq[r][c] = ok ? __half2float(q[((size_t)b * Lq + row) * ldq + h * d + c]) : 0.f;
      sdo[r][c] = ok ? dO[((size_t)b * Lq + row) * lddo + h * d + c] : 0.f;
    }
    if (threadIdx.x < AB_T) {
      const int row = ch * AB_T + threadIdx.x;
      sl[threadIdx.x] = row < Lq ? lse[((size_t)b * nh + h) * Lq + row] : 0.f;
      sD[threadIdx.x] = row < Lq ? Dv[((size_t)b * nh + h) * Lq + row] : 0.f;
    }
    __syncthreads();
    const bool qv = (ch * AB_T + lane) < Lq;   // lane = query within the chunk
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = warp * 4 + j;                // key row owned by this warp
      float s = 0.f, dp = 0.f;
      for (int c = 0; c < d; ++c) {
        s = fmaf(sq[lane][c], sk[r][c], s);
        dp = fmaf(sdo[lane][c], sv[r][c], dp);
      }
      const float p = (qv && (k0 + r) < Lk) ? __expf(s * scale - sl[lane]) : 0.f;
      sp[warp][j][lane] = p;
      sds[warp][j][lane] = p * (dp - sD[lane]) * scale;
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int c = lane + 32 * u;
        if (c < d) {
          float a = ak[j][u], bb = av[j][u];
          for (int qi = 0; qi < AB_T; ++qi) {
            a = fmaf(sds[warp][j][qi], sq[qi][c], a);
            bb = fmaf(sp[warp][j][qi], sdo[qi][c], bb);
          }
          ak[j][u] = a;
          av[j][u] = bb;
        }
      }
    }
    __syncwarp();
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int key = k0 + warp * 4 + j;
    if (key < Lk) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int c = lane + 32 * u;
        if (c < d) {
          dk[((size_t)b * Lk + key) * lddk + h * d + c] = __float2half_rn(ak[j][u]);
          dv[((size_t)b * Lk + key) * lddv + h * d + c] = __float2half_rn(av[j][u]);
        }
      }
    }
  }
}

int attention_bwd_launch(const __half* q, int ldq, const __half* k, int ldk, const __half* v, int ldv, const __half* o,
                         int ldo, const float* dO, int lddo, int B, int heads, int Lq, int Lk, int d, float scale,
                         __half* dq, int lddq, __half* dk, int lddk, __half* dv, int lddv, float* lse_ws, float* d_ws,
                         cudaStream_t stream) {
  if (d > AB_DMAX || d < 1) { set_error("attention_bwd: head dim must be <= 64"); return -1; }
  note("attention_bwd", 10.0 * B * heads * (double)Lq * Lk * d, 0.0);
  DFB_CUDA_OK(launch_pdl(attention_bwd_q_kernel, dim3((Lq + AB_T - 1) / AB_T, heads, B), dim3(256), 0, stream, q, ldq, k,
                         ldk, v, ldv, o, ldo, dO, lddo, Lq, Lk, d, scale, dq, lddq, lse_ws, d_ws));
  if (dk != nullptr && dv != nullptr)
    DFB_CUDA_OK(launch_pdl(attention_bwd_kv_kernel, dim3((Lk + AB_T - 1) / AB_T, heads, B), dim3(256), 0, stream, q, ldq,
                           k, ldk, v, ldv, dO, lddo, Lq, Lk, d, scale, (const float*)lse_ws, (const float*)d_ws, dk,
                           lddk, dv, lddv));
  DFB_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---- GEGLU: proj fp32 [M, 2F] = [value | gate]  ->  h = value * gelu_erf(gate) (fp16), and its backward
__device__ __forceinline__ float gelu_grad_f(float x) {
  // d/dx [ x * Phi(x) ] = Phi(x) + x * phi(x)
  const float phi = 0.3989422804014327f * __expf(-0.5f * x * x);
  const float Phi = 0.5f * (1.f + erff(x * 0.70710678118654752440f));
  return fmaf(x, phi, Phi);
}
__global__ void geglu_fwd_kernel(const float* __restrict__ proj, int F, long total, __half* __restrict__ h) {
  pdl_wait();
  pdl_launch_dependents();
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long m = i / F;
  const int f = (int)(i - m * F);
  const float a = proj[m * 2 * F + f], g = proj[m * 2 * F + F + f];
  h[i] = __float2half_rn(a * 0.5f * g * (1.f + erff(g * 0.70710678118654752440f)));
}
__global__ void geglu_bwd_kernel(const float* __restrict__ proj, const float* __restrict__ dh, int F, long total,
                                 __half* __restrict__ dproj) {
  pdl_wait();
  pdl_launch_dependents();
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long m = i / F;
  const int f = (int)(i - m * F);
  const float a = proj[m * 2 * F + f], g = proj[m * 2 * F + F + f], d = dh[i];
  dproj[m * 2 * F + f] = __float2half_rn(d * 0.5f * g * (1.f + erff(g * 0.70710678118654752440f)));
  dproj[m * 2 * F + F + f] = __float2half_rn(d * a * gelu_grad_f(g));
}
int geglu_fwd_launch(const float* proj, long M, int F, __half* h, cudaStream_t stream) {
  const long total = M * F;
  note("geglu_fwd", 0.0, (double)total * 10.0);
  DFB_CUDA_OK(launch_pdl(geglu_fwd_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, stream, proj, F, total, h));
  DFB_CUDA_OK(cudaGetLastError());
  return 0;
}
int geglu_bwd_launch(const float* proj, const float* dh, long M, int F, __half* dproj, cudaStream_t stream) {
  const long total = M * F;
  note("geglu_bwd", 0.0, (double)total * 16.0);
  DFB_CUDA_OK(launch_pdl(geglu_bwd_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, stream, proj, dh, F, total,
                         dproj));
  DFB_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---- backward-data of the 3x3 / stride 2 / pad 1 Downsample conv (openai_unetmodel.py:134-160): the GEMM
// dcol[mo, (ky*3+kx)*C + c] = sum_co dY[mo, co] W[co, c, ky, kx] is scattered back as a GATHER: input pixel (y, x)
// collects the taps whose output position (y + 1 - ky) / 2, (x + 1 - kx) / 2 is integral and inside.
__global__ void col2im_s2_kernel(const float* __restrict__ dcol, int B, int H, int W, int C, const float* __restrict__ add,
                                 float* __restrict__ dx32, __half* __restrict__ dx16, long total) {
  pdl_wait();
  pdl_launch_dependents();
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  const long pix = i / C;
  const int x = (int)(pix % W), y = (int)((pix / W) % H);
  const long b = pix / ((long)W * H);
  const int Ho = H / 2, Wo = W / 2;
  float s = 0.f;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int ty = y + 1 - ky;
    if (ty < 0 || (ty & 1) || (ty >> 1) >= Ho) continue;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int tx = x + 1 - kx;
      if (tx < 0 || (tx & 1) || (tx >> 1) >= Wo) continue;
      s += dcol[((b * Ho + (ty >> 1)) * Wo + (tx >> 1)) * (9L * C) + (ky * 3 + kx) * C + c];
    }
  }
  if (add != nullptr) s += add[i];
  if (dx32 != nullptr) dx32[i] = s;
  if (dx16 != nullptr) dx16[i] = __float2half_rn(s);
}
int col2im_s2_launch(const float* dcol, int B, int H, int W, int C, const float* add, float* dx32, __half* dx16,
                     cudaStream_t stream) {
  const long total = (long)B * H * W * C;
  note("col2im_s2", 0.0, (double)total * 16.0);
  DFB_CUDA_OK(launch_pdl(col2im_s2_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, stream, dcol, B, H, W, C, add,
                         dx32, dx16, total));
  DFB_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---- classifier head (alignment_backbone.py:676-686 + ddim.py:336-340): c fp32 [B, HW, C] (the last conv's
// output) -> mean over HW -> Linear(C -> 1) -> sigmoid = prob;  seed of d/dc [ log prob ] * seed_scale, fp16:
// dc[b, px, ch] = seed_scale * (1 - prob_b) * w[ch] / HW.   One CTA per sample.
__global__ void __launch_bounds__(256)
classifier_head_kernel(const float* __restrict__ c, int HW, int C, const float* __restrict__ w, const float* __restrict__ bias,
                       float seed_scale, float* __restrict__ prob, __half* __restrict__ dc) {
  __shared__ float red[64];
  __shared__ float sprob;
  pdl_wait();
  pdl_launch_dependents();
  const int b = blockIdx.x;
  const float* cb = c + (size_t)b * HW * C;
  float z = 0.f, dummy = 0.f;
  for (int i = threadIdx.x; i < HW * C; i += blockDim.x) z += cb[i] * w[i % C];
  block_sum2(z, dummy, red);
  if (threadIdx.x == 0) {
    const float p = sigmoid_f(z / (float)HW + bias[0]);
    sprob = p;
    if (prob != nullptr) prob[b] = p;
  }
  __syncthreads();
  if (dc != nullptr) {
    const float f = seed_scale * (1.f - sprob) / (float)HW;
    for (int i = threadIdx.x; i < HW * C; i += blockDim.x) dc[(size_t)b * HW * C + i] = __float2half_rn(f * w[i % C]);
  }
}
int classifier_head_launch(const float* c, int B, int HW, int C, const float* w, const float* bias, float seed_scale,
                           float* prob, __half* dc, cudaStream_t stream) {
  note("classifier_head", 0.0, (double)B * HW * C * 6.0);
  DFB_CUDA_OK(launch_pdl(classifier_head_kernel, dim3(B), dim3(256), 0, stream, c, HW, C, w, bias, seed_scale, prob, dc));
  DFB_CUDA_OK(cudaGetLastError());
  return 0;
}

// y = x * s (fp32), used to undo the gradient's loss scale at the NCHW boundary
__global__ void scale_f32_kernel(float* __restrict__ x, float s, long n) {
  pdl_wait();
  pdl_launch_dependents();
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] *= s;
}
int scale_f32_launch(float* x, float s, long n, cudaStream_t stream) {
  DFB_CUDA_OK(launch_pdl(scale_f32_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, stream, x, s, n));
  DFB_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace dfb
