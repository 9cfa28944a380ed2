// elementwise.cu -- the small bandwidth/latency-bound kernels around the tensor-core GEMMs:
// sinusoidal timestep embedding, fp32->fp16 operand casts (plain, nearest-2x upsample, stride-2
// im2col), the 4-channel stem / head convolutions at the NCHW<->channels-last boundary, and the
// fused classifier-free-guidance + DDIM update.
#include "dfb_internal.h"
#include "dfb_ptx.cuh"

namespace dfb {

// ---------------------------------------------------------------------------------------------
// timestep_embedding (reference util.py:151-171): [cos(t*f_j) | sin(t*f_j)], f_j = exp(-ln(1e4) j/half)
__global__ void temb_kernel(const void* __restrict__ t, int t_is_float, int B, int dim,
                            __half* __restrict__ out) {
  pdl_wait();
  pdl_launch_dependents();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int half_dim = dim / 2;
  if (i >= B * half_dim) return;
  const int b = i / half_dim, j = i - b * half_dim;
  const float tv = t_is_float ? reinterpret_cast<const float*>(t)[b]
                              : (float)reinterpret_cast<const long long*>(t)[b];
  const float freq = expf(-9.210340371976184f * (float)j / (float)half_dim);
  const float arg = tv * freq;
  out[(size_t)b * dim + j] = __float2half_rn(cosf(arg));
  out[(size_t)b * dim + half_dim + j] = __float2half_rn(sinf(arg));
  if ((dim & 1) && j == 0) out[(size_t)b * dim + dim - 1] = __float2half_rn(0.f);
}

int temb_launch(const void* t, int t_is_float, int B, int dim, __half* out, cudaStream_t stream) {
  const int n = B * (dim / 2);
  note("temb", 0.0, (double)B * dim * 2.0);
  DFB_CUDA_OK(launch_pdl(temb_kernel, dim3((n + 255) / 256), dim3(256), 0, stream, t, t_is_float, B, dim, out));
  DFB_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
__global__ void cast_f16_kernel(const float4* __restrict__ src, uint2* __restrict__ dst, size_t n4) {
  pdl_wait();
  pdl_launch_dependents();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n4; i += stride) {
    const float4 v = src[i];
    const __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<const uint32_t*>(&a);
    u.y = *reinterpret_cast<const uint32_t*>(&b);
    dst[i] = u;
  }
}

int cast_f16_launch(const float* src, __half* dst, size_t n, cudaStream_t stream) {
  if (n % 4) {
    set_error("cast_f16: element count must be a multiple of 4");
    return -1;
  }
  const size_t n4 = n / 4;
  const int blocks = (int)std::min<size_t>((n4 + 255) / 256, 148 * 8);
  note("cast_f16", 0.0, (double)n * 6.0);
  DFB_CUDA_OK(launch_pdl(cast_f16_kernel, dim3(blocks), dim3(256), 0, stream, reinterpret_cast<const float4*>(src),
                                              reinterpret_cast<uint2*>(dst), n4));
  DFB_CUDA_OK(cudaGetLastError());
  return 0;
}

// nearest-neighbour 2x upsample (reference openai_unetmodel.py:116, F.interpolate) fused with the
// fp16 cast: dst[b, y, x, :] = src[b, y/2, x/2, :]
__global__ void upsample2x_f16_kernel(const float4* __restrict__ src, uint2* __restrict__ dst, int B,
                                      int H, int W, int C4) {
  pdl_wait();
  pdl_launch_dependents();
  // (32-bit index arithmetic: the launcher rejects tensors of 2^32 elements or more; 64-bit div / mod cost ~100
  // instructions each and were most of this kernel)
  const unsigned total = (unsigned)B * 2u * H * 2u * W * C4;
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned stride = gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    const int c = (int)(i % (unsigned)C4);
    unsigned r = i / (unsigned)C4;
    const int x = (int)(r % (unsigned)(2 * W)); r /= (unsigned)(2 * W);
    const int y = (int)(r % (unsigned)(2 * H));
    const int b = (int)(r / (unsigned)(2 * H));
    const float4 v = src[(((size_t)b * H + (y >> 1)) * W + (x >> 1)) * C4 + c];
    const __half2 a = __floats2half2_rn(v.x, v.y), bb = __floats2half2_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<const uint32_t*>(&a);
    u.y = *reinterpret_cast<const uint32_t*>(&bb);
    dst[i] = u;
  }
}

int upsample2x_f16_launch(const float* src, __half* dst, int B, int H, int W, int C,
                          cudaStream_t stream) {
  if (C % 4) {
    set_error("upsample2x: C must be a multiple of 4");
    return -1;
  }
  const size_t total = (size_t)B * 4 * H * W * (C / 4);
  if (total >= (1ull << 32) - 148ull * 8 * 256) {
    set_error("upsample2x: tensor too large for 32-bit indexing");
    return -1;
  }
  const int blocks = (int)std::min<size_t>((total + 255) / 256, 148 * 8);
  note("upsample2x", 0.0, (double)total * 4 * (2.0 + 1.0));
  DFB_CUDA_OK(launch_pdl(upsample2x_f16_kernel, dim3(blocks), dim3(256), 0, stream, reinterpret_cast<const float4*>(src),
                                                    reinterpret_cast<uint2*>(dst), B, H, W, C / 4));
  DFB_CUDA_OK(cudaGetLastError());
  return 0;
}

// im2col for the three Downsample convs (3x3, stride 2, pad 1; reference openai_unetmodel.py:151):
// dst[(b, yo, xo), tap*C + c] = src[b, 2*yo + dy - 1, 2*xo + dx - 1, c]  (zero outside)
__global__ void im2col_s2_kernel(const float4* __restrict__ src, uint2* __restrict__ dst, int B, int H,
                                 int W, int C4) {
  pdl_wait();
  pdl_launch_dependents();
  const int Ho = H / 2, Wo = W / 2;
  const unsigned total = (unsigned)B * Ho * Wo * 9u * C4;   // (< 2^32: checked by the launcher)
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned stride = gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    const int c = (int)(i % (unsigned)C4);
    unsigned r = i / (unsigned)C4;
    const int tap = (int)(r % 9u); r /= 9u;
    const int xo = (int)(r % (unsigned)Wo); r /= (unsigned)Wo;
    const int yo = (int)(r % (unsigned)Ho);
    const int b = (int)(r / (unsigned)Ho);
    const int y = 2 * yo + tap / 3 - 1, x = 2 * xo + tap % 3 - 1;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (y >= 0 && y < H && x >= 0 && x < W) v = src[(((size_t)b * H + y) * W + x) * C4 + c];
    const __half2 a = __floats2half2_rn(v.x, v.y), bb = __floats2half2_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<const uint32_t*>(&a);
    u.y = *reinterpret_cast<const uint32_t*>(&bb);
    dst[i] = u;
  }
}

int im2col_s2_launch(const float* src, __half* dst, int B, int H, int W, int C, cudaStream_t stream) {
  if (C % 4 || (H & 1) || (W & 1)) {
    set_error("im2col_s2: C must be a multiple of 4 and H, W even");
    return -1;
  }
  const size_t total = (size_t)B * (H / 2) * (W / 2) * 9 * (C / 4);
  if (total >= (1ull << 32) - 148ull * 8 * 256) {
    set_error("im2col_s2: tensor too large for 32-bit indexing");
    return -1;
  }
  const int blocks = (int)std::min<size_t>((total + 255) / 256, 148 * 8);
  note("im2col_s2", 0.0, (double)total * 4 * (2.0 + 1.0));
  DFB_CUDA_OK(launch_pdl(im2col_s2_kernel, dim3(blocks), dim3(256), 0, stream, reinterpret_cast<const float4*>(src),
                                               reinterpret_cast<uint2*>(dst), B, H, W, C / 4));
  DFB_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// stem: conv3x3(Cin->Cout), pad 1, on the NCHW latent (reference openai_unetmodel.py:519).
// w is packed [Cin*9, Cout] (k = ci*9 + tap) so consecutive threads (co) read consecutive floats.
__global__ void stem_conv_kernel(const float* __restrict__ x, int Bsrc, int xoff, int Cin, int H, int W,
                                 const float* __restrict__ w, const float* __restrict__ bias, int Cout,
                                 float* __restrict__ out) {
  pdl_wait();
  pdl_launch_dependents();
  const int pix = blockIdx.x;  // b*H*W + y*W + x
  const int xq = pix % W, yq = (pix / W) % H, b = pix / (W * H);
  const int bs = (b + xoff) % Bsrc;
  extern __shared__ float patch[];  // [Cin*9]
  for (int i = threadIdx.x; i < Cin * 9; i += blockDim.x) {
    const int ci = i / 9, tap = i - ci * 9;
    const int yy = yq + tap / 3 - 1, xx = xq + tap % 3 - 1;
    patch[i] = (yy >= 0 && yy < H && xx >= 0 && xx < W)
                   ? x[(((size_t)bs * Cin + ci) * H + yy) * W + xx]
                   : 0.f;
  }
  __syncthreads();
  for (int co = threadIdx.x; co < Cout; co += blockDim.x) {
    float acc = bias[co];
    for (int i = 0; i < Cin * 9; ++i) acc = fmaf(patch[i], w[(size_t)i * Cout + co], acc);
    out[(size_t)pix * Cout + co] = acc;
  }
}

// The same convolution with one thread per output channel holding its Cin*9 (<= 36) weights in registers and a CTA
// walking STEM_PX pixels whose patches sit in shared memory: the weights are read once per CTA instead of once per
// pixel (the per-pixel kernel above re-read 46 KB of weights from L2 for each of the 2048 pixels and took 27 us of
// the 2.2 ms forward).  Same fma order (bias, then k = ci*9 + tap ascending): bit-identical results.
constexpr int STEM_PX = 16;
constexpr int STEM_K = 36;
__global__ void __launch_bounds__(512)
stem_conv_reg_kernel(const float* __restrict__ x, int Bsrc, int xoff, int Cin, int H, int W, int npix,
                     const float* __restrict__ w, const float* __restrict__ bias, int Cout, float* __restrict__ out) {
  __shared__ __align__(16) float patch[STEM_PX][STEM_K];
  const int co = threadIdx.x, K = Cin * 9;
  float wr[STEM_K];
#pragma unroll
  for (int i = 0; i < STEM_K; ++i) wr[i] = (i < K && co < Cout) ? __ldg(w + (size_t)i * Cout + co) : 0.f;
  const float bv = (co < Cout) ? __ldg(bias + co) : 0.f;
  pdl_wait();   // (weights and bias never depend on the preceding kernel)
  pdl_launch_dependents();
  const int pix0 = blockIdx.x * STEM_PX;
  for (int idx = threadIdx.x; idx < STEM_PX * STEM_K; idx += blockDim.x) {
    const int pp = idx / STEM_K, i = idx - pp * STEM_K;
    const int pix = pix0 + pp;
    float v = 0.f;
    if (pix < npix && i < K) {
      const int xq = pix % W, yq = (pix / W) % H, b = pix / (W * H);
      const int bs = (b + xoff) % Bsrc;
      const int ci = i / 9, tap = i - ci * 9;
      const int yy = yq + tap / 3 - 1, xx = xq + tap % 3 - 1;
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = x[(((size_t)bs * Cin + ci) * H + yy) * W + xx];
    }
    patch[pp][i] = v;
  }
  __syncthreads();
  if (co >= Cout) return;
  // two pixels per pass: the patch rows are read as 9 broadcast float4 each, then two independent fma chains
  // (k >= K: weight and patch are both zero, fma(0, 0, acc) == acc)
#pragma unroll 1
  for (int pp = 0; pp < STEM_PX; pp += 2) {
    if (pix0 + pp >= npix) break;
    float pa[STEM_K], pb[STEM_K];
#pragma unroll
    for (int j = 0; j < STEM_K / 4; ++j) {
      const float4 ta = reinterpret_cast<const float4*>(&patch[pp][0])[j];
      const float4 tb = reinterpret_cast<const float4*>(&patch[pp + 1][0])[j];
      pa[4 * j] = ta.x; pa[4 * j + 1] = ta.y; pa[4 * j + 2] = ta.z; pa[4 * j + 3] = ta.w;
      pb[4 * j] = tb.x; pb[4 * j + 1] = tb.y; pb[4 * j + 2] = tb.z; pb[4 * j + 3] = tb.w;
    }
    float acc0 = bv, acc1 = bv;
#pragma unroll
    for (int i = 0; i < STEM_K; ++i) {
      acc0 = fmaf(pa[i], wr[i], acc0);
      acc1 = fmaf(pb[i], wr[i], acc1);
    }
    out[(size_t)(pix0 + pp) * Cout + co] = acc0;
    if (pix0 + pp + 1 < npix) out[(size_t)(pix0 + pp + 1) * Cout + co] = acc1;
  }
}

int stem_conv_launch(const float* x, int Bsrc, int xoff, int B, int Cin, int H, int W, const float* w,
                     const float* bias, int Cout, float* out, cudaStream_t stream) {
  note("stem_conv", 2.0 * B * H * W * Cin * 9 * Cout, (double)B * H * W * Cout * 4.0);
  if (Cin * 9 <= STEM_K && Cout <= 512) {
    const int npix = B * H * W, threads = (Cout + 31) / 32 * 32;
    DFB_CUDA_OK(launch_pdl(stem_conv_reg_kernel, dim3((npix + STEM_PX - 1) / STEM_PX), dim3(threads), 0, stream, x, Bsrc, xoff,
                           Cin, H, W, npix, w, bias, Cout, out));
    DFB_CUDA_OK(cudaGetLastError());
    return 0;
  }
  DFB_CUDA_OK(launch_pdl(stem_conv_kernel, dim3(B * H * W), dim3(128), Cin * 9 * sizeof(float), stream, x, Bsrc, xoff, Cin, H, W, w, bias,
                                                                        Cout, out));
  DFB_CUDA_OK(cudaGetLastError());
  return 0;
}

// head: conv3x3(C->Cout<=4), pad 1, channels-last fp16 in, NCHW fp32 out (reference
// openai_unetmodel.py:685).  One warp per output pixel, lanes split K = 9*C; w packed [Cout,9,C] fp32.
__global__ void __launch_bounds__(256)
head_conv_kernel(const __half* __restrict__ a, int B, int H, int W, int C, const float* __restrict__ w,
                 const float* __restrict__ bias, int Cout, float* __restrict__ out) {
  pdl_wait();
  pdl_launch_dependents();
  const int pix = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (pix >= B * H * W) return;
  const int xq = pix % W, yq = (pix / W) % H, b = pix / (W * H);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const int c2n = C >> 1;
  for (int tap = 0; tap < 9; ++tap) {
    const int yy = yq + tap / 3 - 1, xx = xq + tap % 3 - 1;
    if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
    const __half2* ar = reinterpret_cast<const __half2*>(a + (((size_t)b * H + yy) * W + xx) * C);
    for (int c2 = lane; c2 < c2n; c2 += 32) {
      const float2 v = __half22float2(ar[c2]);
#pragma unroll
      for (int co = 0; co < 4; ++co) {
        if (co < Cout) {
          const float2 ww = *reinterpret_cast<const float2*>(w + ((size_t)co * 9 + tap) * C + 2 * c2);
          acc[co] = fmaf(v.x, ww.x, acc[co]);
          acc[co] = fmaf(v.y, ww.y, acc[co]);
        }
      }
    }
  }
#pragma unroll
  for (int co = 0; co < 4; ++co) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[co] += __shfl_xor_sync(0xffffffffu, acc[co], o);
  }
  if (lane < Cout) {
    const float r = (lane == 0) ? acc[0] : (lane == 1) ? acc[1] : (lane == 2) ? acc[2] : acc[3];
    out[(((size_t)b * Cout + lane) * H + yq) * W + xq] = r + bias[lane];
  }
}

// The same head with the weights staged once per CTA in shared memory (before the PDL wait: they do not depend on the
// preceding kernel) and HEAD_PPW pixels per warp sharing every weight read; per pixel the lanes split K and
// accumulate in the order of the kernel above, so the results are bit-identical.
constexpr int HEAD_PPW = 2;
template <int NC2>   // NC2 = C / 64: half2 loads per lane and tap
__global__ void __launch_bounds__(256)
head_conv_smem_kernel(const __half* __restrict__ a, int B, int H, int W, int C, const float* __restrict__ w,
                      const float* __restrict__ bias, int Cout, float* __restrict__ out) {
  extern __shared__ float4 head_ws4[];
  const float* ws = reinterpret_cast<const float*>(head_ws4);
  const int n4 = (Cout * 9 * C) >> 2;
  for (int i0 = threadIdx.x; i0 < n4; i0 += 4 * 256) {   // four loads in flight per thread
    float4 t[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) t[j] = (i0 + j * 256 < n4) ? __ldg(reinterpret_cast<const float4*>(w) + i0 + j * 256) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 4; ++j) if (i0 + j * 256 < n4) head_ws4[i0 + j * 256] = t[j];
  }
  __syncthreads();
  pdl_wait();
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  const int pix0 = (blockIdx.x * 8 + (threadIdx.x >> 5)) * HEAD_PPW;
  const int npix = B * H * W;
  if (pix0 >= npix) return;
  // every activation this lane needs (9 taps x NC2 half2 per pixel), issued before the first fma: one round of L2
  // latency instead of one per loop iteration.  Taps outside the image load nothing and contribute fma(0, w, acc).
  __half2 v[HEAD_PPW][9][NC2];
  int xq[HEAD_PPW], yq[HEAD_PPW], bq[HEAD_PPW];
#pragma unroll
  for (int q = 0; q < HEAD_PPW; ++q) {
    const int pix = min(pix0 + q, npix - 1);
    xq[q] = pix % W; yq[q] = (pix / W) % H; bq[q] = pix / (W * H);
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int yy = yq[q] + tap / 3 - 1, xx = xq[q] + tap % 3 - 1;
      const bool ok = (pix0 + q < npix) && yy >= 0 && yy < H && xx >= 0 && xx < W;
      const __half2* ar = reinterpret_cast<const __half2*>(a + (((size_t)bq[q] * H + (ok ? yy : 0)) * W + (ok ? xx : 0)) * C);
#pragma unroll
      for (int k = 0; k < NC2; ++k) v[q][tap][k] = ok ? ar[lane + 32 * k] : __floats2half2_rn(0.f, 0.f);
    }
  }
  float acc[HEAD_PPW][4];
#pragma unroll
  for (int q = 0; q < HEAD_PPW; ++q)
#pragma unroll
    for (int co = 0; co < 4; ++co) acc[q][co] = 0.f;
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
    for (int k = 0; k < NC2; ++k) {
      const int c2 = lane + 32 * k;
      float2 ww[4];
#pragma unroll
      for (int co = 0; co < 4; ++co)
        ww[co] = (co < Cout) ? *reinterpret_cast<const float2*>(ws + ((size_t)co * 9 + tap) * C + 2 * c2) : make_float2(0.f, 0.f);
#pragma unroll
      for (int q = 0; q < HEAD_PPW; ++q) {
        const float2 f = __half22float2(v[q][tap][k]);
#pragma unroll
        for (int co = 0; co < 4; ++co) {
          acc[q][co] = fmaf(f.x, ww[co].x, acc[q][co]);
          acc[q][co] = fmaf(f.y, ww[co].y, acc[q][co]);
        }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < HEAD_PPW; ++q) {
#pragma unroll
    for (int co = 0; co < 4; ++co) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[q][co] += __shfl_xor_sync(0xffffffffu, acc[q][co], o);
    }
    if (lane < Cout && pix0 + q < npix) {
      const float r = (lane == 0) ? acc[q][0] : (lane == 1) ? acc[q][1] : (lane == 2) ? acc[q][2] : acc[q][3];
      out[(((size_t)bq[q] * Cout + lane) * H + yq[q]) * W + xq[q]] = r + bias[lane];
    }
  }
}

int head_conv_launch(const __half* a, int B, int H, int W, int C, const float* w, const float* bias,
                     int Cout, float* out, cudaStream_t stream) {
  if (Cout > 4 || (C & 1)) {
    set_error("head_conv: Cout must be <= 4 and C even");
    return -1;
  }
  note("head_conv", 2.0 * B * H * W * C * 9 * Cout, (double)B * H * W * C * 2.0);
  const size_t wbytes = (size_t)Cout * 9 * C * sizeof(float);
  if (wbytes <= 48 * 1024 && (wbytes & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0) {
    const dim3 grid((B * H * W + 8 * HEAD_PPW - 1) / (8 * HEAD_PPW));
    if (C == 320) {
      DFB_CUDA_OK(launch_pdl(head_conv_smem_kernel<5>, grid, dim3(256), wbytes, stream, a, B, H, W, C, w, bias, Cout, out));
      DFB_CUDA_OK(cudaGetLastError());
      return 0;
    }
    if (C == 128) {
      DFB_CUDA_OK(launch_pdl(head_conv_smem_kernel<2>, grid, dim3(256), wbytes, stream, a, B, H, W, C, w, bias, Cout, out));
      DFB_CUDA_OK(cudaGetLastError());
      return 0;
    }
  }
  DFB_CUDA_OK(launch_pdl(head_conv_kernel, dim3((B * H * W + 7) / 8), dim3(256), 0, stream, a, B, H, W, C, w, bias, Cout, out));
  DFB_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// CAVP helpers (fp16 channels-last).  8 halves (16 B) per thread.
__global__ void im2col_f16_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int NI, int H,
                                  int W, int C8, int kh, int kw, int stride, int pad, int Ho, int Wo,
                                  int K8) {
  pdl_wait();
  pdl_launch_dependents();
  const size_t total = (size_t)NI * Ho * Wo * K8;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t step = (size_t)gridDim.x * blockDim.x;
  for (; i < total; i += step) {
    const int k8 = (int)(i % K8);
    size_t r = i / K8;
    const int xo = (int)(r % Wo); r /= Wo;
    const int yo = (int)(r % Ho);
    const int n = (int)(r / Ho);
    uint4 v = make_uint4(0, 0, 0, 0);
    const int tap = k8 / C8, c8 = k8 - tap * C8;
    if (tap < kh * kw) {
      const int y = yo * stride + tap / kw - pad, x = xo * stride + tap % kw - pad;
      if (y >= 0 && y < H && x >= 0 && x < W) v = src[(((size_t)n * H + y) * W + x) * C8 + c8];
    }
    dst[i] = v;
  }
}

int im2col_f16_launch(const __half* src, __half* dst, int NI, int H, int W, int C, int kh, int kw,
                      int stride, int pad, int Kpad, cudaStream_t stream) {
  if (C % 8 || Kpad % 8 || Kpad < kh * kw * C) {
    set_error("im2col_f16: C and Kpad must be multiples of 8 and Kpad >= kh*kw*C");
    return -1;
  }
  const int Ho = (H + 2 * pad - kh) / stride + 1, Wo = (W + 2 * pad - kw) / stride + 1;
  const size_t total = (size_t)NI * Ho * Wo * (Kpad / 8);
  const int blocks = (int)std::min<size_t>((total + 255) / 256, 148 * 16);
  note("im2col_f16", 0.0, (double)total * 32.0);
  DFB_CUDA_OK(launch_pdl(im2col_f16_kernel, dim3(blocks), dim3(256), 0, stream,
                         reinterpret_cast<const uint4*>(src), reinterpret_cast<uint4*>(dst), NI, H, W, C / 8,
                         kh, kw, stride, pad, Ho, Wo, Kpad / 8));
  return 0;
}

__global__ void pool2d_f16_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int NI, int H,
                                  int W, int C8, int kh, int kw, int sh, int sw, int ph, int pw, int Ho,
                                  int Wo, int is_max) {
  pdl_wait();
  pdl_launch_dependents();
  const size_t total = (size_t)NI * Ho * Wo * C8;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t step = (size_t)gridDim.x * blockDim.x;
  for (; i < total; i += step) {
    const int c8 = (int)(i % C8);
    size_t r = i / C8;
    const int xo = (int)(r % Wo); r /= Wo;
    const int yo = (int)(r % Ho);
    const int n = (int)(r / Ho);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = is_max ? -INFINITY : 0.f;
    for (int a = 0; a < kh; ++a)
      for (int b = 0; b < kw; ++b) {
        const int y = yo * sh + a - ph, x = xo * sw + b - pw;
        if (y < 0 || y >= H || x < 0 || x >= W) continue;  // max: padding ignored; avg: windows never pad here
        const uint4 u = src[(((size_t)n * H + y) * W + x) * C8 + c8];
        const __half2* hp = reinterpret_cast<const __half2*>(&u);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __half22float2(hp[j]);
          if (is_max) { acc[2 * j] = fmaxf(acc[2 * j], f.x); acc[2 * j + 1] = fmaxf(acc[2 * j + 1], f.y); }
          else { acc[2 * j] += f.x; acc[2 * j + 1] += f.y; }
        }
      }
    const float sc = is_max ? 1.f : 1.f / (float)(kh * kw);
    uint4 o;
    __half2* op = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) op[j] = __floats2half2_rn(acc[2 * j] * sc, acc[2 * j + 1] * sc);
    dst[i] = o;
  }
}

int pool2d_f16_launch(const __half* src, __half* dst, int NI, int H, int W, int C, int kh, int kw,
                      int sh, int sw, int ph, int pw, int is_max, cudaStream_t stream) {
  if (C % 8) {
    set_error("pool2d_f16: C must be a multiple of 8");
    return -1;
  }
  const int Ho = (H + 2 * ph - kh) / sh + 1, Wo = (W + 2 * pw - kw) / sw + 1;
  const size_t total = (size_t)NI * Ho * Wo * (C / 8);
  const int blocks = (int)std::min<size_t>((total + 255) / 256, 148 * 16);
  note("pool2d_f16", 0.0, (double)total * 16.0 * (kh * kw + 1));
  DFB_CUDA_OK(launch_pdl(pool2d_f16_kernel, dim3(blocks), dim3(256), 0, stream,
                         reinterpret_cast<const uint4*>(src), reinterpret_cast<uint4*>(dst), NI, H, W, C / 8,
                         kh, kw, sh, sw, ph, pw, Ho, Wo, is_max));
  return 0;
}

// ---------------------------------------------------------------------------------------------
// One DDIM step's arithmetic (reference ddim.py:241-245 CFG, :377-380 classifier term, :258-273
// update), same operation order as the reference and no FMA contraction so the sampler arithmetic
// itself is bit-compatible with the fp32 host path:
//   e      = e_u + s (e_c - e_u)            [ - grad_coef * grad ]
//   x0     = (x - sqrt(1-a_t) e) / sqrt(a_t)
//   x_prev = sqrt(a_prev) x0 + sqrt(1 - a_prev - sigma^2) e          (eta = 0: no noise term)
__global__ void ddim_update_kernel(const float* __restrict__ x, const float* __restrict__ eu,
                                   const float* __restrict__ ec, const float* __restrict__ grad,
                                   float s, float sqrt_1mat, float sqrt_at, float sqrt_aprev,
                                   float dir_coef, float grad_coef, float* __restrict__ x_prev,
                                   float* __restrict__ pred_x0, size_t n) {
  pdl_wait();
  pdl_launch_dependents();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float e;
  if (eu != nullptr)
    e = __fadd_rn(eu[i], __fmul_rn(s, __fsub_rn(ec[i], eu[i])));
  else
    e = ec[i];
  if (grad != nullptr) e = __fsub_rn(e, __fmul_rn(grad_coef, grad[i]));
  const float x0 = __fdiv_rn(__fsub_rn(x[i], __fmul_rn(sqrt_1mat, e)), sqrt_at);
  const float xp = __fadd_rn(__fmul_rn(sqrt_aprev, x0), __fmul_rn(dir_coef, e));
  x_prev[i] = xp;
  if (pred_x0 != nullptr) pred_x0[i] = x0;
}

int ddim_update_launch(const float* x, const float* eps_uncond, const float* eps_cond,
                       const float* grad, float cfg_scale, float sqrt_one_minus_at, float sqrt_at,
                       float sqrt_a_prev, float dir_coef, float grad_coef, float* x_prev,
                       float* pred_x0, size_t n, cudaStream_t stream) {
  note("ddim_update", 0.0, (double)n * 20.0);
  DFB_CUDA_OK(launch_pdl(ddim_update_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, stream, 
      x, eps_uncond, eps_cond, grad, cfg_scale, sqrt_one_minus_at, sqrt_at, sqrt_a_prev, dir_coef,
      grad_coef, x_prev, pred_x0, n));
  DFB_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---- CAVP frame ingest (row N4; reference inference/demo_util.py:135-163: per frame cv2 BGR->RGB, PIL
// Resize((224,224)) = Pillow's two-pass antialiased bilinear resample in 8-bit fixed point, ToTensor).
// Pillow's arithmetic (src/libImaging/Resample.c: ImagingResampleHorizontal_8bpc / Vertical_8bpc): per output
// sample  clip8((2^21 + sum_x pix[xmin + x] * k[x]) >> 22)  with int32 coefficients k = round(w * 2^22), the
// horizontal pass rounded to uint8 before the vertical one.  The host builds the coefficient tables exactly as
// precompute_coeffs / normalize_coeffs_8bpc do (diff_foley_b200/frames.py); these kernels do the integer MACs,
// so the result is bit-identical to PIL for every frame size.  A whole window of frames is one launch pair.
__device__ __forceinline__ int clip8_fixed(int v) {
  v >>= 22;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}
// in [N,H,W,3] u8 -> tmp [N,H,OW,3] u8 (channel order swapped when swap_rb: cv2 delivers BGR)
__global__ void resample_h_u8_kernel(const uint8_t* __restrict__ in, int H, int W, int OW, int swap_rb,
                                     const int* __restrict__ kk, const int* __restrict__ bounds, int ksize,
                                     uint8_t* __restrict__ tmp, long total) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;   // (n*H + y) * OW + ox
  if (i >= total) return;
  const int ox = (int)(i % OW);
  const long row = i / OW;
  const int xmin = bounds[2 * ox], xmax = bounds[2 * ox + 1];
  const int* k = kk + (long)ox * ksize;
  const uint8_t* src = in + (row * W + xmin) * 3;
  int s0 = 1 << 21, s1 = 1 << 21, s2 = 1 << 21;
  for (int x = 0; x < xmax; ++x) {
    const int c = k[x];
    s0 += src[3 * x + 0] * c;
    s1 += src[3 * x + 1] * c;
    s2 += src[3 * x + 2] * c;
  }
  uint8_t* d = tmp + i * 3;
  d[swap_rb ? 2 : 0] = (uint8_t)clip8_fixed(s0);
  d[1] = (uint8_t)clip8_fixed(s1);
  d[swap_rb ? 0 : 2] = (uint8_t)clip8_fixed(s2);
}
// tmp [N,H,OW,3] u8 -> out [N,3,OH,OW] fp32 in [0,1] (ToTensor: x / 255), optionally the uint8 image too
__global__ void resample_v_u8_kernel(const uint8_t* __restrict__ tmp, int H, int OW, int OH,
                                     const int* __restrict__ kk, const int* __restrict__ bounds, int ksize,
                                     float* __restrict__ out, uint8_t* __restrict__ out_u8, long total) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;   // (n*OH + oy) * OW + ox
  if (i >= total) return;
  const int ox = (int)(i % OW);
  const int oy = (int)((i / OW) % OH);
  const long n = i / ((long)OW * OH);
  const int ymin = bounds[2 * oy], ymax = bounds[2 * oy + 1];
  const int* k = kk + (long)oy * ksize;
  const uint8_t* src = tmp + ((n * H + ymin) * OW + ox) * 3;
  int s0 = 1 << 21, s1 = 1 << 21, s2 = 1 << 21;
  for (int y = 0; y < ymax; ++y) {
    const int c = k[y];
    const uint8_t* q = src + (long)y * OW * 3;
    s0 += q[0] * c;
    s1 += q[1] * c;
    s2 += q[2] * c;
  }
  const int v0 = clip8_fixed(s0), v1 = clip8_fixed(s1), v2 = clip8_fixed(s2);
  const long plane = (long)OH * OW, o = n * 3 * plane + (long)oy * OW + ox;
  out[o] = __fdiv_rn((float)v0, 255.f);
  out[o + plane] = __fdiv_rn((float)v1, 255.f);
  out[o + 2 * plane] = __fdiv_rn((float)v2, 255.f);
  if (out_u8 != nullptr) { out_u8[i * 3] = (uint8_t)v0; out_u8[i * 3 + 1] = (uint8_t)v1; out_u8[i * 3 + 2] = (uint8_t)v2; }
}

int frames_resize_launch(const uint8_t* src, int N, int H, int W, int swap_rb, const int* kk_h, const int* bounds_h,
                         int ksize_h, int OW, const int* kk_v, const int* bounds_v, int ksize_v, int OH, uint8_t* tmp,
                         float* out, uint8_t* out_u8, cudaStream_t stream) {
  const long t1 = (long)N * H * OW, t2 = (long)N * OH * OW;
  note("frames_resize", 0.0, (double)N * H * W * 3 + 2.0 * t1 * 3 + t2 * 12.0);
  DFB_CUDA_OK(launch_pdl(resample_h_u8_kernel, dim3((unsigned)((t1 + 255) / 256)), dim3(256), 0, stream, src, H, W, OW,
                         swap_rb, kk_h, bounds_h, ksize_h, tmp, t1));
  DFB_CUDA_OK(launch_pdl(resample_v_u8_kernel, dim3((unsigned)((t2 + 255) / 256)), dim3(256), 0, stream,
                         (const uint8_t*)tmp, H, OW, OH, kk_v, bounds_v, ksize_v, out, out_u8, t2));
  DFB_CUDA_OK(cudaGetLastError());
  return 0;
}

int elementwise_init() {
#define DFB_MAXSHARED(k) \
  DFB_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared))
  DFB_MAXSHARED(temb_kernel);
  DFB_MAXSHARED(cast_f16_kernel);
  DFB_MAXSHARED(upsample2x_f16_kernel);
  DFB_MAXSHARED(im2col_s2_kernel);
  DFB_MAXSHARED(stem_conv_kernel);
  DFB_MAXSHARED(head_conv_kernel);
  DFB_MAXSHARED(head_conv_smem_kernel<5>);
  DFB_MAXSHARED(head_conv_smem_kernel<2>);
  DFB_MAXSHARED(im2col_f16_kernel);
  DFB_MAXSHARED(pool2d_f16_kernel);
  DFB_MAXSHARED(ddim_update_kernel);
#undef DFB_MAXSHARED
  return 0;
}

}  // namespace dfb
