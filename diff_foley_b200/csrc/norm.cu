// norm.cu -- GroupNorm(32)+SiLU and LayerNorm producers of the fp16 GEMM operands.
//
// Both read the fp32 residual stream (channels-last) and emit the fp16 tile the following tensor-core
// GEMM / conv pulls in through TMA, so the normalised activation is written exactly once, at half
// width.  Statistics are fp32: LayerNorm two-pass (mean, then centred second moment) on registers,
// GroupNorm one pass about a pivot (no cancellation, one reduction round).
//
// Replaces: GroupNorm32 + SiLU (reference util.py:214-216, openai_unetmodel.py:201-203,225-228,
// 682-684), Normalize (attention_openai.py:76-77, eps 1e-6), nn.LayerNorm
// (attention_openai.py:203-205) and the th.cat of the skip connection (openai_unetmodel.py:736):
// the kernel normalises across the *concatenation* of two sources without materialising it in fp32.
#include <algorithm>
#include <cstdlib>

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
constexpr int GN_MAXP = 8;   // float2 pairs held per thread

// grid = (P, 32 groups, B) with cluster dims (P,1,1): the P CTAs of a cluster share one (sample,
// group) slab, split along the pixel axis.  Each thread keeps its <= 8 channel pairs in registers (one
// global read, no shared-memory slab); the group statistics are one pass about a pivot (see below), the
// per-CTA partials exchanged through distributed shared memory.
__global__ void __launch_bounds__(GN_THREADS)
groupnorm_kernel(const float* __restrict__ src0, int C0, const float* __restrict__ src1, int C1,
                 int HW, const float* __restrict__ gamma, const float* __restrict__ beta,
                 float eps, int silu, __half* __restrict__ out, __half* __restrict__ raw_out,
                 unsigned long long* trace) {
  __shared__ float red[GN_THREADS / 32], red2[GN_THREADS / 32];
  __shared__ __align__(8) float2 inbox[8];   // (s1, s2) of every CTA of the cluster, pushed by its owner
  __shared__ __align__(8) uint64_t inbox_bar;
  if (threadIdx.x == 0) trace_mark(trace, 0);
  if (gridDim.x > 1) {
    // partial sums are exchanged by st.async pushes counted on the receiver's mbarrier: no cluster-scope
    // fence (a release/acquire cluster barrier costs a MEMBAR.ALL.GPU, > 1 us here).  The relaxed
    // arrive/wait pair only proves that every peer's barrier exists; its latency hides under the loads.
    if (threadIdx.x == 0) {
      mbar_init(&inbox_bar, 1);
      fence_mbar_init();
      mbar_expect_tx(&inbox_bar, gridDim.x * 8u);
    }
    cluster_arrive_relaxed();
  }
  pdl_wait();
  if (threadIdx.x == 0) trace_mark(trace, 1);
  pdl_launch_dependents();
  const int P = gridDim.x, rank = blockIdx.x;
  const int C = C0 + C1;
  const int cpg = C / 32;
  const int hp = cpg >> 1;  // channel pairs per pixel
  const int g = blockIdx.y, b = blockIdx.z;
  const int cbase = g * cpg;
  const int px_per = (HW + P - 1) / P;
  const int px0 = rank * px_per;
  const int npx = max(0, min(px_per, HW - px0));
  const int npairs = npx * hp;

  // One-pass statistics about a pivot (the slab's first element, the same for every CTA of the cluster):
  // s1 = sum(x - piv), s2 = sum((x - piv)^2).  The pivot is a sample of the data, so |mean - piv| is of
  // the order of the standard deviation and var = s2/n - (s1/n)^2 loses no precision to cancellation --
  // and one block reduction + one cluster exchange replace the two of a mean-then-variance scheme
  // (this kernel is a pure latency chain: load, reduce, exchange, write).
  const float piv = (cbase < C0) ? __ldg(src0 + (size_t)b * HW * C0 + cbase)
                                 : __ldg(src1 + (size_t)b * HW * C1 + (cbase - C0));
  float2 v[GN_MAXP];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int k = 0; k < GN_MAXP; ++k) {
    const int i = threadIdx.x + k * GN_THREADS;
    v[k] = make_float2(0.f, 0.f);
    if (i < npairs) {
      const int px = px0 + i / hp;
      const int c = cbase + 2 * (i % hp);
      v[k] = (c < C0) ? *reinterpret_cast<const float2*>(src0 + ((size_t)b * HW + px) * C0 + c)
                      : *reinterpret_cast<const float2*>(src1 + ((size_t)b * HW + px) * C1 + (c - C0));
    }
  }
#pragma unroll
  for (int k = 0; k < GN_MAXP; ++k) {
    if (threadIdx.x + k * GN_THREADS < npairs) {
      const float dx = v[k].x - piv, dy = v[k].y - piv;
      s1 += dx + dy;
      s2 += dx * dx + dy * dy;
    }
  }
  const float inv_n = 1.f / (float)(HW * cpg);
  {
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { red[warp] = s1; red2[warp] = s2; }
    __syncthreads();
    s1 = warp_sum(lane < GN_THREADS / 32 ? red[lane] : 0.f);
    s2 = warp_sum(lane < GN_THREADS / 32 ? red2[lane] : 0.f);
    if (P > 1) {
      cluster_wait();
      if (threadIdx.x < P)
        st_async_f2(mapa_shared(smem_u32(&inbox[rank]), threadIdx.x), s1, s2,
                    mapa_shared(smem_u32(&inbox_bar), threadIdx.x));
      mbar_wait_cluster(&inbox_bar, 0);
      s1 = 0.f; s2 = 0.f;
      for (int r = 0; r < P; ++r) { s1 += inbox[r].x; s2 += inbox[r].y; }  // rank order: deterministic
    }
  }
  const float dm = s1 * inv_n;
  const float mean = piv + dm;
  const float var = fmaxf(s2 * inv_n - dm * dm, 0.f);
  const float rstd = rsqrtf(var + eps);
#pragma unroll
  for (int k = 0; k < GN_MAXP; ++k) {
    const int i = threadIdx.x + k * GN_THREADS;
    if (i < npairs) {
      const int px = px0 + i / hp;
      const int c = cbase + 2 * (i % hp);
      float y0 = (v[k].x - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
      float y1 = (v[k].y - mean) * rstd * __ldg(gamma + c + 1) + __ldg(beta + c + 1);
      if (silu) {
        y0 = y0 / (1.f + expf(-y0));
        y1 = y1 / (1.f + expf(-y1));
      }
      const size_t o = ((size_t)b * HW + px) * C + c;
      *reinterpret_cast<__half2*>(out + o) = __floats2half2_rn(y0, y1);
      if (raw_out != nullptr) *reinterpret_cast<__half2*>(raw_out + o) = __floats2half2_rn(v[k].x, v[k].y);
    }
  }
  if (threadIdx.x == 0) trace_mark(trace, 7);
}


// ---- v2 (default): pixel-major CTAs.  The kernel above gives every (sample, group) its own cluster, so
// a thread's loads are 8 bytes out of a 40-160 byte run per pixel and the per-element index math (runtime
// div / mod) dominates the instruction stream.  Here a CTA owns CW consecutive channels (a whole number of
// groups, >= 128 bytes per pixel) x a contiguous run of pixels; thread (ty, tx) keeps channel pair tx of
// pixels ty, ty + NY, ... in registers, so a warp's loads are contiguous runs, the channel -- hence the
// group, gamma, beta, source tensor -- of a thread never changes, and there is no division in any loop.
// grid = (P pixel chunks, C / CW channel chunks, B), cluster (P,1,1).  Statistics: per-thread partials about
// the group's pivot -> shared memory, one warp per group sums them in a fixed order -> st.async push of
// the group's (s1, s2) to every CTA of the cluster (as in v1) -> every thread sums the P slots of its own
// group in rank order.  Deterministic (no atomics).
constexpr int GN2_THREADS = 512;
template <int KMAX>
__global__ void __launch_bounds__(GN2_THREADS)
groupnorm2_kernel(const float* __restrict__ src0, int C0, const float* __restrict__ src1, int C1, int HW,
                  const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int silu,
                  __half* __restrict__ out, __half* __restrict__ raw_out, int CW, int NY, int px_per,
                  unsigned long long* trace) {
  __shared__ float part1[GN2_THREADS], part2[GN2_THREADS];
  __shared__ __align__(8) float2 inbox[8 * 8];   // [rank][local group] (s1, s2), pushed by the owners
  __shared__ __align__(8) uint64_t inbox_bar;
  if (threadIdx.x == 0) trace_mark(trace, 0);
  const int P = gridDim.x, rank = blockIdx.x;
  const int C = C0 + C1, cpg = C >> 5, hp = cpg >> 1;
  const int TX = CW >> 1, ngl = CW / cpg;
  if (P > 1) {
    if (threadIdx.x == 0) {
      mbar_init(&inbox_bar, 1);
      fence_mbar_init();
      mbar_expect_tx(&inbox_bar, (uint32_t)(P * ngl) * 8u);
    }
    cluster_arrive_relaxed();
  }
  const int b = blockIdx.z;
  const int tid = threadIdx.x;
  const int ty = tid / TX, tx = tid - ty * TX;
  const bool act = ty < NY;
  const int c = blockIdx.y * CW + 2 * tx;       // this thread's channel pair (never straddles a group)
  const int gl = (2 * tx) / cpg;                  // local group
  const int cg0 = blockIdx.y * CW + gl * cpg;     // first channel of the group: the pivot's channel
  const float2 gm = __ldg(reinterpret_cast<const float2*>(gamma + c));
  const float2 bt = __ldg(reinterpret_cast<const float2*>(beta + c));
  const int px0 = rank * px_per;
  const int px1 = min(px0 + px_per, HW);
  pdl_wait();
  if (threadIdx.x == 0) trace_mark(trace, 1);
  pdl_launch_dependents();
  const float* base;
  int pstride;
  if (c < C0) { base = src0 + (size_t)b * HW * C0 + c; pstride = C0; }
  else { base = src1 + (size_t)b * HW * C1 + (c - C0); pstride = C1; }
  const float piv = (cg0 < C0) ? __ldg(src0 + (size_t)b * HW * C0 + cg0)
                               : __ldg(src1 + (size_t)b * HW * C1 + (cg0 - C0));
  float2 v[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    const int px = px0 + ty + k * NY;
    v[k] = make_float2(piv, piv);   // contributes 0 to both sums
    if (act && px < px1) v[k] = *reinterpret_cast<const float2*>(base + (size_t)px * pstride);
  }
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    const float dx = v[k].x - piv, dy = v[k].y - piv;
    s1 += dx + dy;
    s2 = fmaf(dx, dx, fmaf(dy, dy, s2));
  }
  // group-contiguous slot: [local group][ty][pair within group]
  const int E = hp * NY;                        // partials per group
  if (act) {
    const int slot = gl * E + ty * hp + (tx - gl * hp);
    part1[slot] = s1;
    part2[slot] = s2;
  }
  __syncthreads();
  if (P > 1) cluster_wait();   // every peer's inbox barrier exists (arrive was issued at kernel entry)
  const int warp = tid >> 5, lane = tid & 31;
  for (int g = warp; g < ngl; g += GN2_THREADS / 32) {
    float a1 = 0.f, a2 = 0.f;
    for (int e = lane; e < E; e += 32) { a1 += part1[g * E + e]; a2 += part2[g * E + e]; }
    a1 = warp_sum(a1);
    a2 = warp_sum(a2);
    if (P == 1) {
      if (lane == 0) inbox[g] = make_float2(a1, a2);
    } else {
      if (lane < P)
        st_async_f2(mapa_shared(smem_u32(&inbox[rank * 8 + g]), lane), a1, a2,
                    mapa_shared(smem_u32(&inbox_bar), lane));
    }
  }
  if (P > 1) {
    mbar_wait_cluster(&inbox_bar, 0);
  } else {
    __syncthreads();
  }
  float t1 = 0.f, t2 = 0.f;
  for (int r = 0; r < P; ++r) { const float2 q = inbox[r * 8 + gl]; t1 += q.x; t2 += q.y; }  // rank order
  const float inv_n = 1.f / (float)(HW * cpg);
  const float dm = t1 * inv_n;
  const float mean = piv + dm;
  const float rstd = rsqrtf(fmaxf(t2 * inv_n - dm * dm, 0.f) + eps);
  const float a0 = rstd * gm.x, a1 = rstd * gm.y;
  const float b0 = fmaf(-mean, a0, bt.x), b1 = fmaf(-mean, a1, bt.y);
  __half* obase = out + (size_t)b * HW * C + c;
  __half* rbase = raw_out ? raw_out + (size_t)b * HW * C + c : nullptr;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    const int px = px0 + ty + k * NY;
    if (act && px < px1) {
      float y0 = fmaf(v[k].x, a0, b0), y1 = fmaf(v[k].y, a1, b1);
      if (silu) { y0 = silu_f(y0); y1 = silu_f(y1); }
      *reinterpret_cast<__half2*>(obase + (size_t)px * C) = __floats2half2_rn(y0, y1);
      if (rbase) *reinterpret_cast<__half2*>(rbase + (size_t)px * C) = __floats2half2_rn(v[k].x, v[k].y);
    }
  }
  if (threadIdx.x == 0) trace_mark(trace, 7);
}

// ---- slabs too large for one cluster's registers (the first-stage decoder: 128..512 channels at up to
// 128x512 pixels, model.py:557-663 of the reference's stage1_autoencoder).  Two kernels over pixel
// chunks: (1) every CTA reduces its chunk about the slab's pivot and writes (s1, s2) to a slot of a small
// workspace -- no atomics, the slots are summed in order by (2), which re-reads its chunk, normalises and
// writes fp16.  grid = (chunks, 32 groups, B).
constexpr int GNB_THREADS = 256;
constexpr int GNB_CHUNK_PX = 1024;  // pixels per CTA
static float* g_gn_ws = nullptr;    // [B][32][chunks][2]
constexpr size_t GN_WS_FLOATS = 1 << 18;

__device__ __forceinline__ const float* gn_src(const float* src0, int C0, const float* src1, int C1, size_t pix,
                                               int c) {
  return (c < C0) ? src0 + pix * C0 + c : src1 + pix * C1 + (c - C0);
}

__global__ void __launch_bounds__(GNB_THREADS)
groupnorm_big_stats_kernel(const float* __restrict__ src0, int C0, const float* __restrict__ src1, int C1,
                           int HW, float* __restrict__ ws) {
  __shared__ float red[GNB_THREADS / 32], red2[GNB_THREADS / 32];
  pdl_wait();
  pdl_launch_dependents();
  const int C = C0 + C1, cpg = C / 32, hp = cpg >> 1;
  const int chunk = blockIdx.x, g = blockIdx.y, b = blockIdx.z, cbase = g * cpg;
  const int px0 = chunk * GNB_CHUNK_PX, npx = min(GNB_CHUNK_PX, HW - px0);
  const float piv = *gn_src(src0, C0, src1, C1, (size_t)b * HW, cbase);
  float s1 = 0.f, s2 = 0.f;
  for (int i = threadIdx.x; i < npx * hp; i += GNB_THREADS) {
    const int px = px0 + i / hp, c = cbase + 2 * (i % hp);
    const float2 v = *reinterpret_cast<const float2*>(gn_src(src0, C0, src1, C1, (size_t)b * HW + px, c));
    const float dx = v.x - piv, dy = v.y - piv;
    s1 += dx + dy;
    s2 += dx * dx + dy * dy;
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[warp] = s1; red2[warp] = s2; }
  __syncthreads();
  if (warp == 0) {
    s1 = warp_sum(lane < GNB_THREADS / 32 ? red[lane] : 0.f);
    s2 = warp_sum(lane < GNB_THREADS / 32 ? red2[lane] : 0.f);
    if (lane == 0) {
      float* slot = ws + (((size_t)b * 32 + g) * gridDim.x + chunk) * 2;
      slot[0] = s1;
      slot[1] = s2;
    }
  }
}

__global__ void __launch_bounds__(GNB_THREADS)
groupnorm_big_apply_kernel(const float* __restrict__ src0, int C0, const float* __restrict__ src1, int C1,
                           int HW, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                           int silu, const float* __restrict__ ws, __half* __restrict__ out,
                           __half* __restrict__ raw_out) {
  pdl_wait();
  pdl_launch_dependents();
  const int C = C0 + C1, cpg = C / 32, hp = cpg >> 1;
  const int chunk = blockIdx.x, g = blockIdx.y, b = blockIdx.z, cbase = g * cpg;
  const int px0 = chunk * GNB_CHUNK_PX, npx = min(GNB_CHUNK_PX, HW - px0);
  const float piv = *gn_src(src0, C0, src1, C1, (size_t)b * HW, cbase);
  float s1 = 0.f, s2 = 0.f;
  const float* slots = ws + ((size_t)b * 32 + g) * gridDim.x * 2;
  for (int k = 0; k < (int)gridDim.x; ++k) { s1 += slots[2 * k]; s2 += slots[2 * k + 1]; }  // in order
  const float inv_n = 1.f / ((float)HW * (float)cpg);
  const float dm = s1 * inv_n, mean = piv + dm;
  const float rstd = rsqrtf(fmaxf(s2 * inv_n - dm * dm, 0.f) + eps);
  for (int i = threadIdx.x; i < npx * hp; i += GNB_THREADS) {
    const int px = px0 + i / hp, c = cbase + 2 * (i % hp);
    const float2 v = *reinterpret_cast<const float2*>(gn_src(src0, C0, src1, C1, (size_t)b * HW + px, c));
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

static int groupnorm_big_launch(const float* src0, int C0, const float* src1, int C1, int B, int HW,
                                const float* gamma, const float* beta, float eps, int silu, __half* out,
                                __half* raw_out, cudaStream_t stream) {
  const int chunks = (HW + GNB_CHUNK_PX - 1) / GNB_CHUNK_PX;
  if ((size_t)B * 32 * chunks * 2 > GN_WS_FLOATS) {
    set_error("groupnorm: batch x pixels too large for the statistics workspace");
    return -1;
  }
  if (g_gn_ws == nullptr) DFB_CUDA_OK(cudaMalloc(&g_gn_ws, GN_WS_FLOATS * sizeof(float)));
  const int C = C0 + C1;
  note("groupnorm", 0.0, (double)B * HW * C * (8.0 + 2.0 + (raw_out ? 2.0 : 0.0)), B * HW, C, 0, 1, 32 * B * chunks);
  const dim3 grid(chunks, 32, B);
  DFB_CUDA_OK(launch_pdl(groupnorm_big_stats_kernel, grid, dim3(GNB_THREADS), 0, stream, src0, C0, src1, C1, HW,
                         g_gn_ws));
  DFB_CUDA_OK(launch_pdl(groupnorm_big_apply_kernel, grid, dim3(GNB_THREADS), 0, stream, src0, C0, src1, C1, HW,
                         gamma, beta, eps, silu, (const float*)g_gn_ws, out, raw_out));
  DFB_CUDA_OK(cudaGetLastError());
  return 0;
}

// large-batch variants of the norm launchers (on by default, DFB_NORM_STREAM=0 disables): see
// layernorm_stream_kernel and the cluster-size policy in groupnorm_launch; +3.8 % at B = 8 clips
static bool norm_stream_enabled() {
  static const int v = getenv("DFB_NORM_STREAM") ? atoi(getenv("DFB_NORM_STREAM")) : 1;
  return v != 0;
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
  static const int gn_ver = getenv("DFB_GN") ? atoi(getenv("DFB_GN")) : 2;
  if (gn_ver == 2) {
    // channels per CTA: a whole number of groups, >= 32 channels (128-byte runs per pixel)
    const int cpg = C / 32;
    int CW = cpg * ((32 + cpg - 1) / cpg);
    while (C % CW) CW += cpg;
    const int TX = CW / 2;
    const int chunks = C / CW;
    // pixel chunks (= cluster size): aim at >= ~128 CTAs, at least 16 pixels per CTA, registers <= 16 pairs
    int P = 1;
    static const int gn_target = getenv("DFB_GN_CTAS") ? atoi(getenv("DFB_GN_CTAS")) : 128;
    while (P < 8 && (long)chunks * B * P < gn_target && HW / (P * 2) >= 16) P *= 2;
    int NYmax = GN2_THREADS / TX;
    while (P < 8 && (HW + P - 1) / P > NYmax * 16) P *= 2;
    const int px_per = (HW + P - 1) / P;
    if (px_per <= NYmax * 16 && CW / cpg <= 8) {
      const int NY = std::min(NYmax, px_per);
      const int kmax = (px_per + NY - 1) / NY;
      note("groupnorm", 0.0, (double)B * HW * C * (4.0 + 2.0 + (raw_out ? 2.0 : 0.0)), B * HW, C, 0, 1, chunks * B * P);
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(P, chunks, B);
      cfg.blockDim = dim3(GN2_THREADS);
      cfg.stream = stream;
      cudaLaunchAttribute attr[2];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
      attr[1].id = cudaLaunchAttributeClusterDimension;
      attr[1].val.clusterDim.x = P;
      attr[1].val.clusterDim.y = 1;
      attr[1].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 2;
      auto go = [&](auto kern) {
        return cudaLaunchKernelEx(&cfg, kern, src0, C0, src1, C1, HW, gamma, beta, eps, silu, out, raw_out, CW, NY,
                                  px_per, trace_record());
      };
      if (kmax <= 4) DFB_CUDA_OK(go(groupnorm2_kernel<4>));
      else if (kmax <= 8) DFB_CUDA_OK(go(groupnorm2_kernel<8>));
      else DFB_CUDA_OK(go(groupnorm2_kernel<16>));
      return 0;
    }
  }
  const long pairs = (long)HW * (C / 64);
  const long cap = (long)GN_THREADS * GN_MAXP;
  int P = (int)((pairs + cap - 1) / cap);
  // spread a slab over a few SMs when there are enough pixels to split (latency, not capacity)
  const bool grid_full = norm_stream_enabled() && 32L * B * std::max(P, 1) >= 296;  // >= 2 CTAs per SM anyway
  if (!grid_full) { if (HW >= 1024) P = std::max(P, 4); else if (HW >= 256) P = std::max(P, 2); }
  if (P > 8)  // slab beyond one cluster's registers: statistics + apply kernels over pixel chunks
    return groupnorm_big_launch(src0, C0, src1, C1, B, HW, gamma, beta, eps, silu, out, raw_out, stream);
  note("groupnorm", 0.0, (double)B * HW * C * (4.0 + 2.0 + (raw_out ? 2.0 : 0.0)), B * HW, C, 0, 1, 32 * B * P);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(P, 32, B);
  cfg.blockDim = dim3(GN_THREADS);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  attr[1].id = cudaLaunchAttributeClusterDimension;
  attr[1].val.clusterDim.x = P;
  attr[1].val.clusterDim.y = 1;
  attr[1].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  DFB_CUDA_OK(cudaLaunchKernelEx(&cfg, groupnorm_kernel, src0, C0, src1, C1, HW, gamma, beta, eps, silu, out,
                                 raw_out, trace_record()));
  return 0;
}

// One warp per row of fp32 [rows, C].  The row lives in registers (<= 10 float4 per lane, C <= 1280):
// one global read, two-pass statistics on the registers, fp16 write.  4 warps per CTA so that even the
// 128-row deep levels spread over 32 SMs; up to 16 for the 2048-row level so that the grid stays at
// ~128 CTAs (the in-kernel timeline showed the dependent launch being released ~2 us later behind a
// 512-CTA grid than behind a 128-CTA one).
constexpr int LN_WARPS = 4;
constexpr int LN_MAX_WARPS = 16;
constexpr int LN_MAXV = 10;
__global__ void __launch_bounds__(LN_MAX_WARPS * 32)
layernorm_kernel(const float* __restrict__ src, int rows, int C, const float* __restrict__ gamma,
                 const float* __restrict__ beta, float eps, __half* __restrict__ out,
                 unsigned long long* trace) {
  if (threadIdx.x == 0) trace_mark(trace, 0);
  pdl_wait();
  if (threadIdx.x == 0) trace_mark(trace, 1);
  pdl_launch_dependents();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* x = reinterpret_cast<const float4*>(src + (size_t)row * C);
  const int n4 = C >> 2;
  float4 v[LN_MAXV];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < LN_MAXV; ++k) {
    const int i = lane + 32 * k;
    v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n4) {
      v[k] = x[i];
      s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
    }
  }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < LN_MAXV; ++k) {
    if (lane + 32 * k < n4) {
      const float a = v[k].x - mean, b = v[k].y - mean, c = v[k].z - mean, d = v[k].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
  __half* o = out + (size_t)row * C;
#pragma unroll
  for (int k = 0; k < LN_MAXV; ++k) {
    const int i = lane + 32 * k;
    if (i < n4) {
      const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + i);
      const float4 bt = __ldg(reinterpret_cast<const float4*>(beta) + i);
      const __half2 h0 = __floats2half2_rn((v[k].x - mean) * rstd * gm.x + bt.x,
                                           (v[k].y - mean) * rstd * gm.y + bt.y);
      const __half2 h1 = __floats2half2_rn((v[k].z - mean) * rstd * gm.z + bt.z,
                                           (v[k].w - mean) * rstd * gm.w + bt.w);
      uint2 u;
      u.x = *reinterpret_cast<const uint32_t*>(&h0);
      u.y = *reinterpret_cast<const uint32_t*>(&h1);
      *reinterpret_cast<uint2*>(o + 4 * i) = u;
    }
  }
  if (lane == 0) trace_mark(trace, 7);
}

// Large row counts (B_eff >= 16 at the 16x64 / 8x32 levels): the one-row-per-warp kernel above is bound
// by how few bytes each SM has in flight (one 1.3-2.5 KB row per warp, then the warp retires).  Here a
// warp streams over rows two at a time -- both rows' loads issued before either reduction -- with gamma
// and beta held in registers; V = float4 per lane per row (C <= 128 V).
template <int V>
__global__ void __launch_bounds__(256)
layernorm_stream_kernel(const float* __restrict__ src, int rows, int C, const float* __restrict__ gamma,
                        const float* __restrict__ beta, float eps, __half* __restrict__ out,
                        unsigned long long* trace) {
  if (threadIdx.x == 0) trace_mark(trace, 0);
  const int lane = threadIdx.x & 31;
  const int n4 = C >> 2;
  float4 gm[V], bt[V];
#pragma unroll
  for (int k = 0; k < V; ++k) {
    const int i = lane + 32 * k;
    gm[k] = bt[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n4) {
      gm[k] = __ldg(reinterpret_cast<const float4*>(gamma) + i);
      bt[k] = __ldg(reinterpret_cast<const float4*>(beta) + i);
    }
  }
  pdl_wait();
  if (threadIdx.x == 0) trace_mark(trace, 1);
  pdl_launch_dependents();
  const int wg = blockIdx.x * 8 + (threadIdx.x >> 5), nw = gridDim.x * 8;
  const float inv_c = 1.f / (float)C;
  for (int row = 2 * wg; row < rows; row += 2 * nw) {
    const bool two = row + 1 < rows;
    const float4* xa = reinterpret_cast<const float4*>(src + (size_t)row * C);
    const float4* xb = xa + (two ? n4 : 0);
    float4 a[V], b[V];
    float sa = 0.f, sb = 0.f;
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const int i = lane + 32 * k;
      a[k] = b[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < n4) { a[k] = xa[i]; b[k] = xb[i]; }
    }
#pragma unroll
    for (int k = 0; k < V; ++k) {
      sa += (a[k].x + a[k].y) + (a[k].z + a[k].w);
      sb += (b[k].x + b[k].y) + (b[k].z + b[k].w);
    }
    const float ma = warp_sum(sa) * inv_c, mb = warp_sum(sb) * inv_c;
    float qa = 0.f, qb = 0.f;
#pragma unroll
    for (int k = 0; k < V; ++k) {
      if (lane + 32 * k < n4) {
        float d0 = a[k].x - ma, d1 = a[k].y - ma, d2 = a[k].z - ma, d3 = a[k].w - ma;
        qa += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
        d0 = b[k].x - mb; d1 = b[k].y - mb; d2 = b[k].z - mb; d3 = b[k].w - mb;
        qb += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
      }
    }
    const float ra = rsqrtf(warp_sum(qa) * inv_c + eps), rb = rsqrtf(warp_sum(qb) * inv_c + eps);
    __half* oa = out + (size_t)row * C;
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const int i = lane + 32 * k;
      if (i < n4) {
        __half2 h0 = __floats2half2_rn((a[k].x - ma) * ra * gm[k].x + bt[k].x, (a[k].y - ma) * ra * gm[k].y + bt[k].y);
        __half2 h1 = __floats2half2_rn((a[k].z - ma) * ra * gm[k].z + bt[k].z, (a[k].w - ma) * ra * gm[k].w + bt[k].w);
        uint2 u;
        u.x = *reinterpret_cast<const uint32_t*>(&h0);
        u.y = *reinterpret_cast<const uint32_t*>(&h1);
        *reinterpret_cast<uint2*>(oa + 4 * i) = u;
        if (two) {
          h0 = __floats2half2_rn((b[k].x - mb) * rb * gm[k].x + bt[k].x, (b[k].y - mb) * rb * gm[k].y + bt[k].y);
          h1 = __floats2half2_rn((b[k].z - mb) * rb * gm[k].z + bt[k].z, (b[k].w - mb) * rb * gm[k].w + bt[k].w);
          u.x = *reinterpret_cast<const uint32_t*>(&h0);
          u.y = *reinterpret_cast<const uint32_t*>(&h1);
          *reinterpret_cast<uint2*>(oa + C + 4 * i) = u;
        }
      }
    }
  }
  if (threadIdx.x == 0) trace_mark(trace, 7);
}

int layernorm_launch(const float* src, int rows, int C, const float* gamma, const float* beta,
                     float eps, __half* out, cudaStream_t stream) {
  if (C % 4 != 0 || C > 128 * LN_MAXV) {
    set_error("layernorm: C must be a multiple of 4 and <= 1280");
    return -1;
  }
  if (norm_stream_enabled() && rows >= 4096 && C <= 640) {
    const int nblk = std::min((rows / 2 + 7) / 8, 148 * 8);
    note("layernorm", 0.0, (double)rows * C * 6.0, rows, C, 0, 1, nblk);
    if (C <= 384)
      DFB_CUDA_OK(launch_pdl(layernorm_stream_kernel<3>, dim3(nblk), dim3(256), 0, stream, src, rows, C, gamma, beta,
                             eps, out, trace_record()));
    else
      DFB_CUDA_OK(launch_pdl(layernorm_stream_kernel<5>, dim3(nblk), dim3(256), 0, stream, src, rows, C, gamma, beta,
                             eps, out, trace_record()));
    DFB_CUDA_OK(cudaGetLastError());
    return 0;
  }
  const int warps = std::min(LN_MAX_WARPS, std::max(LN_WARPS, rows / 128));
  const int nblk = (rows + warps - 1) / warps;
  note("layernorm", 0.0, (double)rows * C * 6.0, rows, C, 0, 1, nblk);
  DFB_CUDA_OK(launch_pdl(layernorm_kernel, dim3(nblk), dim3(warps * 32), 0, stream, src, rows, C, gamma, beta, eps, out,
                         trace_record()));
  DFB_CUDA_OK(cudaGetLastError());
  return 0;
}

// softmax(scale * x) over the rows of fp32 [rows, n] -> fp16 (the first-stage decoder's single-head
// 512-wide attention, model.py:245-300: its head dim exceeds the fused attention kernel's TMEM budget, so
// q k^T and P v run as plain GEMMs around this kernel).  One warp per row, the row in registers.
constexpr int SM_MAXV = 16;  // float4 per lane: n <= 2048
__global__ void __launch_bounds__(128)
softmax_rows_kernel(const float* __restrict__ src, int rows, int n, float scale, __half* __restrict__ out) {
  pdl_wait();
  pdl_launch_dependents();
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* x = reinterpret_cast<const float4*>(src + (size_t)row * n);
  const int n4 = n >> 2;
  float4 v[SM_MAXV];
  float mx = -INFINITY;
#pragma unroll
  for (int k = 0; k < SM_MAXV; ++k) {
    const int i = lane + 32 * k;
    if (i < n4) {
      v[k] = x[i];
      mx = fmaxf(mx, fmaxf(fmaxf(v[k].x, v[k].y), fmaxf(v[k].z, v[k].w)));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  const float sl = scale * 1.4426950408889634f, ml = mx * sl;
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < SM_MAXV; ++k) {
    if (lane + 32 * k < n4) {
      v[k].x = exp2f(fmaf(v[k].x, sl, -ml)); v[k].y = exp2f(fmaf(v[k].y, sl, -ml));
      v[k].z = exp2f(fmaf(v[k].z, sl, -ml)); v[k].w = exp2f(fmaf(v[k].w, sl, -ml));
      sum += (v[k].x + v[k].y) + (v[k].z + v[k].w);
    }
  }
  const float inv = 1.f / warp_sum(sum);
  __half* o = out + (size_t)row * n;
#pragma unroll
  for (int k = 0; k < SM_MAXV; ++k) {
    const int i = lane + 32 * k;
    if (i < n4) {
      const __half2 h0 = __floats2half2_rn(v[k].x * inv, v[k].y * inv);
      const __half2 h1 = __floats2half2_rn(v[k].z * inv, v[k].w * inv);
      uint2 u;
      u.x = *reinterpret_cast<const uint32_t*>(&h0);
      u.y = *reinterpret_cast<const uint32_t*>(&h1);
      *reinterpret_cast<uint2*>(o + 4 * i) = u;
    }
  }
}

int softmax_rows_launch(const float* src, int rows, int n, float scale, __half* out, cudaStream_t stream) {
  if (n % 4 != 0 || n > 128 * SM_MAXV || rows < 1) {
    set_error("softmax_rows: n must be a multiple of 4 and <= 2048");
    return -1;
  }
  note("softmax", 0.0, (double)rows * n * 6.0, rows, n, 0, 1, (rows + 3) / 4);
  DFB_CUDA_OK(launch_pdl(softmax_rows_kernel, dim3((rows + 3) / 4), dim3(128), 0, stream, src, rows, n, scale, out));
  DFB_CUDA_OK(cudaGetLastError());
  return 0;
}

// All kernels of the plan ask for the same (maximum-shared) L1/shared-memory split: the GEMM and
// attention kernels need ~200 KB of shared memory, and an SM has to drain to change its carve-out, so a
// small-smem kernel with the default preference between two GEMMs would force two reconfigurations.
int norm_init() {
  DFB_CUDA_OK(cudaFuncSetAttribute(groupnorm_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                   cudaSharedmemCarveoutMaxShared));
  DFB_CUDA_OK(cudaFuncSetAttribute(layernorm_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                   cudaSharedmemCarveoutMaxShared));
  DFB_CUDA_OK(cudaFuncSetAttribute(groupnorm2_kernel<4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  DFB_CUDA_OK(cudaFuncSetAttribute(groupnorm2_kernel<8>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  DFB_CUDA_OK(cudaFuncSetAttribute(groupnorm2_kernel<16>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  return 0;
}

}  // namespace dfb
