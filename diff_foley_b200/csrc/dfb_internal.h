// dfb_internal.h -- host-side declarations shared by the kernels' launchers, the UNet engine and
// the C ABI (include/dfb.h).  Not installed; the public surface is include/dfb.h only.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <utility>

namespace dfb {

// ---------------------------------------------------------------------------- error plumbing
void set_error(const std::string& msg);
const char* last_error();
#define DFB_CUDA_OK(expr)                                                                   \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      (void)cudaGetLastError(); /* clear the non-sticky error state */                       \
      ::dfb::set_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " at " + \
                       __FILE__ + ":" + std::to_string(__LINE__));                          \
      return -2;                                                                            \
    }                                                                                       \
  } while (0)

// Every launcher records what it just enqueued (kind + algorithmic flops / bytes) so the profiler
// entry point (dfb_unet_profile) can attribute event-timed durations without a second op table.
struct LaunchNote {
  const char* kind;
  double flops, bytes;
  int M, N, K, splits, ctas;
};
extern thread_local LaunchNote g_note;
inline void note(const char* kind, double flops, double bytes, int M = 0, int N = 0, int K = 0,
                 int splits = 1, int ctas = 0) {
  g_note.kind = kind; g_note.flops = flops; g_note.bytes = bytes;
  g_note.M = M; g_note.N = N; g_note.K = K; g_note.splits = splits; g_note.ctas = ctas;
}

// In-kernel timeline (dfb_unet_trace): while buf != nullptr the instrumented launchers hand their
// kernel the 32-word record of op `next`; the engine sets `next` before each op of the plan.
struct TraceState {
  unsigned long long* buf = nullptr;
  int cap = 0, next = 0;
};
extern TraceState g_trace;
inline unsigned long long* trace_record() {
  if (g_trace.buf == nullptr || g_trace.next >= g_trace.cap) return nullptr;
  return g_trace.buf + 32ull * g_trace.next;
}

// Launch with the programmatic-dependent-launch attribute (DFB_NO_PDL=1 disables it).
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                              cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// one-time per-process kernel attribute setup (dynamic shared memory opt-in); idempotent
int kernels_init();
int igemm_init();
int norm_init();
int attention_init();
int elementwise_init();

// ------------------------------------------------------------------- implicit-GEMM (tcgen05)
// One kernel covers Linear, 1x1 conv, 3x3 conv (9 shifted TMA boxes, zero-filled halo) and the
// CAVP (3,1,1) temporal conv: the A operand is an fp16 channels-last activation tensor
// [B,T,H,W,C] read through a 5-D TMA map, the weights are fp16 [N, taps*C] (K-major).
enum : int { ACT_NONE = 0, ACT_SILU = 1, ACT_GEGLU = 2, ACT_RELU = 3 };

struct IGemmGeom {
  // activation tensor dims (output spatial dims == input dims; stride-1 "same" convs only)
  int B, T, H, W, C;
  // tile box: bb*bt*bh*bw == 128 output positions per M tile
  int bb, bt, bh, bw;
  int ntaps;            // 1, 3 (temporal) or 9 (3x3)
  int8_t dt[9], dh[9], dw[9];
  // optional second A source: an fp16 channels-last tensor [B,T,H,W,C2] over the same positions whose C2
  // channels are appended to K after the taps (tap offset 0) -- a 1x1 conv on another tensor summed into
  // the same accumulator (the ResBlock skip connection fused into conv2); weights [N, ntaps*C + C2]
  int C2;
  const __half* A2;
};

struct IGemmEpilogue {
  float* out_f32;           // optional [M, ldo]
  __half* out_f16;          // optional [M, ldo]
  int ldo;                  // row stride (elements) of the outputs
  const float* bias;        // optional [N]
  const float* rowvec;      // optional per-sample vector [B, ld_rowvec] added to every row of sample b
  int ld_rowvec;
  int rows_per_sample;      // rows (output positions) per sample, for rowvec
  const int* rowvec_row;    // optional device scalar: use row *rowvec_row of `rowvec` for EVERY sample (the
                            // sampler's table of per-step timestep embeddings, indexed by its step counter)
  const float* residual;    // optional fp32 [M, ld_res]
  const __half* residual_f16;  // optional fp16 [M, ld_res] (CAVP ResNet identity path)
  int ld_res;
  int act;                  // ACT_*
  // LayerNorm folded into the GEMM (attention_openai.py:211-215: x + f(norm(x)) sub-blocks).  Consumer:
  // A = raw fp16 x, weights pre-multiplied by gamma, `bias` = t_n = sum_k beta_k W[n,k] (+ the layer's
  // bias), ln_s[n] = sum_k W'[n,k]; out = rstd_m * acc - rstd_m * mu_m * ln_s[n] + t_n with the row
  // statistics merged from the producer's per-(row, N-tile) partials.  Producer: stats_out [M, tiles_n].
  const float2* ln_stats;
  int ln_tiles;
  float ln_inv_c, ln_eps;
  const float* ln_s;
  float2* stats_out;
};

struct IGemmPlan {
  CUtensorMap tmA, tmW, tmA2;
  IGemmGeom g;
  IGemmEpilogue e;
  int M, N, K;       // logical sizes: M = B*T*H*W, K = ntaps*C
  int BN;            // tile N: 64, 128, or 256 = the wide variant (bn_run columns of it computed)
  int bn_run = 0;    // columns per tile actually computed (== BN unless BN == 256)
  int tiles_m, tiles_n, splits;  // splits == cluster size along grid.z
  int deep = -1;     // operand ring: 1 = deep (1 CTA/SM), 0 = shallow (2 CTAs/SM), -1 = launcher's default
  int pair = 0;      // 1 = CTA-pair tiles (tcgen05.mma.cta_group::2, 256 x BN per SM pair), no split-K
  int a32 = 0;       // 1 = 32-row activation box + 9-stage ring (<= 32 output positions, weight streaming)
  // weights of the NEXT GEMM of the plan: every CTA issues an L2 prefetch for a slice of them, so the
  // next (weight-streaming) kernel finds its operand in L2 instead of waiting on cold HBM misses
  const void* next_w = nullptr;
  size_t next_w_bytes = 0;
};

// Builds tensor maps + tiling for `A` (fp16 [B,T,H,W,C]) and `Wt` (fp16 [N, ntaps*C]).
// splits==0 lets the planner pick tile width and split-K factor (= cluster size, <= 8) from its cost
// model; split-K partials are reduced through distributed shared memory, no global workspace.
int igemm_plan(IGemmPlan* plan, const __half* A, const __half* Wt, int N, const IGemmGeom& g,
               const IGemmEpilogue& e, int splits);
int igemm_launch(const IGemmPlan& plan, cudaStream_t stream);
// tuning aid (tools/autotune_igemm.py): force the tile width (64 / 128, 0 = planner's choice) and the
// ring depth (1 / 0, -1 = default) of every plan built afterwards
void igemm_force(int bn, int deep);
// CTA-pair (cta_group::2) tiles for every plan built afterwards that can take them: 1 on, 0 off, -1 = DFB_PAIR env
void igemm_force_pair(int pair);
// geometry for a "same"-padded stride-1 conv with a (kt,kh,kw) kernel over [B,T,H,W,C] (kt*kh*kw <= 9)
IGemmGeom conv_taps_geom(int B, int T, int H, int W, int C, int kt, int kh, int kw);
// convenience geometry for a plain [M,K] x [N,K]^T GEMM
IGemmGeom gemm_geom(int M, int K);
// geometry for a 3x3 / pad 1 / stride 1 conv over [B,H,W,C]
IGemmGeom conv3x3_geom(int B, int H, int W, int C);

// --------------------------------------------------------------------------------- norm kernels
// GroupNorm(32 groups) over the channel-concatenation of up to two fp32 NHWC sources, optional
// SiLU, fp16 output [B,HW,C0+C1]; optionally also the raw (un-normalised) fp16 copy.
int groupnorm_launch(const float* src0, int C0, const float* src1, int C1, int B, int HW,
                     const float* gamma, const float* beta, float eps, int silu, __half* out,
                     __half* raw_out, cudaStream_t stream);
// LayerNorm over the last dim of fp32 [rows, C] -> fp16.
int layernorm_launch(const float* src, int rows, int C, const float* gamma, const float* beta,
                     float eps, __half* out, cudaStream_t stream);

// softmax(scale * x) over the rows of fp32 [rows, n] -> fp16 (n <= 2048)
int softmax_rows_launch(const float* src, int rows, int n, float scale, __half* out, cudaStream_t stream);

// ---------------------------------------------------------------------------------- attention
// q: fp16 rows [B*Lq, ldq] (head h at columns h*dpad), k/v likewise with Lk rows per sample.
// out: fp16 [B*Lq, ldo], head h at columns h*d (un-padded).
int attention_launch(const __half* q, int ldq, const __half* k, int ldk, const __half* v, int ldv,
                     __half* out, int ldo, int B, int heads, int Lq, int Lk, int d, int dpad,
                     float scale, cudaStream_t stream);

// -------------------------------------------------------------------------------- elementwise
int temb_launch(const void* t, int t_is_float, int B, int dim, __half* out, cudaStream_t stream);
int cast_f16_launch(const float* src, __half* dst, size_t n, cudaStream_t stream);
int upsample2x_f16_launch(const float* src, __half* dst, int B, int H, int W, int C,
                          cudaStream_t stream);
// im2col for 3x3 / pad 1 / stride 2: fp32 NHWC [B,H,W,C] -> fp16 [B*(H/2)*(W/2), 9*C]
int im2col_s2_launch(const float* src, __half* dst, int B, int H, int W, int C,
                     cudaStream_t stream);
// stem conv: x NCHW fp32 [Bsrc,Cin,H,W] (sample b reads x[(b + xoff) % Bsrc]) -> NHWC fp32 [B,H,W,Cout]
int stem_conv_launch(const float* x, int Bsrc, int xoff, int B, int Cin, int H, int W, const float* w,
                     const float* bias, int Cout, float* out, cudaStream_t stream);
// head conv: fp16 NHWC [B,H,W,C] (already GN+SiLU'd) -> NCHW fp32 [B,Cout,H,W]
int head_conv_launch(const __half* a, int B, int H, int W, int C, const float* w,
                     const float* bias, int Cout, float* out, cudaStream_t stream);
// ---- fp16 channels-last helpers used by the CAVP encoders
// generic im2col: [NI,H,W,C] -> [NI*Ho*Wo, Kpad], k = (ky*kw+kx)*C + c, zero beyond kh*kw*C
int im2col_f16_launch(const __half* src, __half* dst, int NI, int H, int W, int C, int kh, int kw,
                      int stride, int pad, int Kpad, cudaStream_t stream);
// max / average pooling over [NI,H,W,C]; window (kh,kw), stride (sh,sw), padding (ph,pw)
int pool2d_f16_launch(const __half* src, __half* dst, int NI, int H, int W, int C, int kh, int kw,
                      int sh, int sw, int ph, int pw, int is_max, cudaStream_t stream);
// CAVP frame ingest: Pillow-exact two-pass 8-bit bilinear resample of N uint8 HWC frames + ToTensor
int frames_resize_launch(const uint8_t* src, int N, int H, int W, int swap_rb, const int* kk_h, const int* bounds_h,
                         int ksize_h, int OW, const int* kk_v, const int* bounds_v, int ksize_v, int OH, uint8_t* tmp,
                         float* out, uint8_t* out_u8, cudaStream_t stream);
// fused classifier-free-guidance combine + DDIM update (ddim.py:241-273 of the reference)
int ddim_update_launch(const float* x, const float* eps_uncond, const float* eps_cond,
                       const float* grad, float cfg_scale, float sqrt_one_minus_at, float sqrt_at,
                       float sqrt_a_prev, float dir_coef, float grad_coef, float* x_prev,
                       float* pred_x0, size_t n, cudaStream_t stream);

// ------------------------------------------------------------- backward kernels (backward.cu)
int groupnorm_bwd_launch(const float* x, int C, int B, int HW, const float* gamma, const float* beta, float eps,
                         int silu, const float* dy, const float* add, float* dx32, __half* dx16, cudaStream_t stream);
int layernorm_bwd_launch(const float* x, int rows, int C, const float* gamma, float eps, const float* dy,
                         const float* add, float* dx32, __half* dx16, cudaStream_t stream);
int attention_bwd_launch(const __half* q, int ldq, const __half* k, int ldk, const __half* v, int ldv, const __half* o,
                         int ldo, const float* dO, int lddo, int B, int heads, int Lq, int Lk, int d, float scale,
                         __half* dq, int lddq, __half* dk, int lddk, __half* dv, int lddv, float* lse_ws, float* d_ws,
                         cudaStream_t stream);
int geglu_fwd_launch(const float* proj, long M, int F, __half* h, cudaStream_t stream);
int geglu_bwd_launch(const float* proj, const float* dh, long M, int F, __half* dproj, cudaStream_t stream);
int col2im_s2_launch(const float* dcol, int B, int H, int W, int C, const float* add, float* dx32, __half* dx16,
                     cudaStream_t stream);
int classifier_head_launch(const float* c, int B, int HW, int C, const float* w, const float* bias, float seed_scale,
                           float* prob, __half* dc, cudaStream_t stream);
int scale_f32_launch(float* x, float s, long n, cudaStream_t stream);

}  // namespace dfb
