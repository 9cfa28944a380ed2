/* dfb.h -- C ABI of libdfb.so, the B200 (sm_100a) implementation of Diff-Foley's latent-diffusion
 * sampling hot path.  Plain pointers and sizes only; every pointer named *_dev is a CUDA device
 * pointer owned by the caller, `stream` is a cudaStream_t passed as void*.  Every function returns
 * 0 on success and a negative DFB_E* code on failure; dfb_last_error() returns a thread-local
 * message.  No C++ exceptions cross this boundary and there is no CPU fallback: on a host without an
 * sm_100 device the compute entry points fail with DFB_E_CUDA.
 *
 * Citations are file:line in the reference repository (luosiallen/Diff-Foley @ 0ba1e8ad).
 */
#ifndef DFB_H_
#define DFB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DFB_OK 0
#define DFB_E_INVALID (-1) /* bad argument / unsupported configuration */
#define DFB_E_CUDA (-2)    /* CUDA runtime error, message has details   */
#define DFB_E_DRIVER (-3)  /* tensor-map encode / driver entry point    */
#define DFB_E_STATE (-4)   /* call order (e.g. forward before finalize) */

#define DFB_ACT_NONE 0
#define DFB_ACT_SILU 1
#define DFB_ACT_GEGLU 2
#define DFB_ACT_RELU 3

typedef struct dfb_unet* dfb_handle;

/* Mirrors the constructor arguments of UNetModel that the inference configs use
 * (diff_foley/modules/diffusionmodules/openai_unetmodel.py:443-469, inference/config/
 * Stage2_LDM.yaml:21-36).  Unsupported combinations are rejected by dfb_unet_create. */
typedef struct dfb_unet_cfg {
  int32_t in_channels;
  int32_t model_channels;
  int32_t out_channels;
  int32_t num_res_blocks;
  int32_t n_channel_mult;
  int32_t channel_mult[8];
  int32_t n_attention_resolutions;
  int32_t attention_resolutions[8];
  int32_t num_heads;
  int32_t context_dim;
  int32_t latent_h; /* 16 : mel bins / 8   (ddpm.py:1293) */
  int32_t latent_w; /* 64 : mel frames / 8                 */
  int32_t max_context_len; /* 32 CAVP frames (notebook cell 13) */
  int32_t max_batch;       /* largest B_eff (= 2 x clips under classifier-free guidance) */
} dfb_unet_cfg;

const char* dfb_last_error(void);
const char* dfb_version(void);

/* ------------------------------------------------------------------ UNet engine (rows a1-a12)
 * Replaces UNetModel.forward (openai_unetmodel.py:710-742) and everything below it. */
int dfb_unet_create(const dfb_unet_cfg* cfg, int device, dfb_handle* out);
/* `name` is the reference state-dict key relative to the UNet (e.g. "input_blocks.1.0.in_layers.2.weight",
 * demo_util.py:182-184 / SURVEY 3.4); `src_dev` is fp32 on the device in the reference's own layout.
 * The engine packs it into its fp16 tensor-core layout immediately; the caller may free src after. */
int dfb_unet_set_weight(dfb_handle h, const char* name, const float* src_dev, const int64_t* shape,
                        int ndim);
/* Checks every parameter was supplied and builds the per-batch launch plans lazily. */
int dfb_unet_finalize(dfb_handle h);
/* Number of parameter tensors the engine expects; names via dfb_unet_weight_name(i). */
int dfb_unet_num_weights(dfb_handle h);
const char* dfb_unet_weight_name(dfb_handle h, int i);
/* Pre-computes the (step-invariant) cross-attention K/V of all 16 transformer blocks from the
 * cross-attention context [B_eff, ctx_len, context_dim] fp32 (attention_openai.py:175-176). */
int dfb_unet_set_context(dfb_handle h, const float* ctx_dev, int b_eff, int ctx_len, void* stream);
/* eps = UNet(x, t, context).  x/out: fp32 NCHW [b_eff, C, H, W].  t: int64 or fp32 [b_eff].
 * ctx_dev may be NULL to reuse the K/V of the last dfb_unet_set_context call.
 * If x has only b_eff/x_repeat distinct samples (CFG feeds cat([x, x])), pass x_repeat = 2 and the
 * un-duplicated tensor. */
int dfb_unet_forward(dfb_handle h, const float* x_dev, int x_repeat, const void* t_dev,
                     int t_is_float, const float* ctx_dev, int ctx_len, float* out_dev, int b_eff,
                     void* stream);
/* Whole DDIM loop with classifier-free guidance (ddim.py:179-273): x_T -> x_0 in place.
 * cond/uncond: fp32 [B, ctx_len, context_dim]; n_steps = the number of entries of the schedule arrays (for
 * the 'uniform' discretisation that is len(range(0, 1000, 1000 // S)), which exceeds S when S does not
 * divide 1000 -- the reference runs them all, ddim.py:197-199); timesteps: host int64[n_steps] in sampling
 * order (961, 921, ... 1); coefficient arrays are host fp32[n_steps] in the same order:
 * sqrt(1-a_t), sqrt(a_t), sqrt(a_prev), sqrt(1-a_prev-sigma^2).  The step is captured once as a CUDA
 * graph over engine-owned staging buffers and replayed n_steps times.  pred_x0_dev may be NULL;
 * x_first_dev / pred_first_dev (optional) receive x and pred_x0 after the FIRST step, which the
 * reference's sampler logs into its intermediates (ddim.py:223-226).
 * With a communicator (dfb_comm_init) the 2B (clip, branch) units are sharded over the ranks: every rank
 * passes the same (replicated) x / cond / uncond, evaluates its slice of units, and the per-step eps
 * all-gather (ncclAllGather, captured inside the step graph) is the only exchange. */
int dfb_ddim_sample(dfb_handle h, float* x_dev, const float* cond_dev, const float* uncond_dev,
                    int n_clips, int ctx_len, float cfg_scale, int n_steps, const int64_t* timesteps,
                    const float* sqrt_one_minus_at, const float* sqrt_at, const float* sqrt_a_prev,
                    const float* dir_coef, float* pred_x0_dev, float* x_first_dev, float* pred_first_dev,
                    void* stream);
/* DPM-Solver++(2M) -- the reference's default sampler in the notebook (dpm_solver/sampler.py:89-156:
 * predict_x0, multistep order 2, 'time_uniform', lower_order_final) -- with classifier-free guidance, same
 * fused step graph as dfb_ddim_sample.  Host arrays of n_evals entries, one per model evaluation k (at the
 * continuous time t_k): t_input[k] = fractional model-input time (t_k - 1/N) * 1000 (dpm_solver.py:278-287),
 * sigma[k], alpha[k] = marginal std / mean coefficient at t_k (data prediction, :386-394), cx[k] =
 * sigma_{k+1}/sigma_k, a_coef[k] = alpha_{k+1} (e^{-h} - 1), inv_r0[k] = h / h_0, order[k] in {1, 2}
 * (:504-533, :755-790).  n_evals <= 256.  x: x_T -> x_0 in place; pred_x0_dev (optional) = last data prediction. */
int dfb_dpm_solver_sample(dfb_handle h, float* x_dev, const float* cond_dev, const float* uncond_dev,
                          int n_clips, int ctx_len, float cfg_scale, int n_evals, const float* t_input,
                          const float* sigma, const float* alpha, const float* cx, const float* a_coef,
                          const float* inv_r0, const int32_t* order, float* pred_x0_dev, void* stream);
/* ---- backward ops of the double-guidance classifier gradient (ddim.py:333-341 through Classifier_Backbone,
 * alignment_backbone.py:417-686).  Backward-data of convs / Linears are dfb_conv3x3 / dfb_gemm with rotated /
 * transposed weights; these are the remaining pieces.  fp32 channels-last activations [B,HW,C]; `add` (optional
 * fp32) is summed into dx (residual branch); dx is written as fp32 and/or fp16 (either may be NULL). */
int dfb_groupnorm_bwd(const float* x_dev, int C, int B, int HW, const float* gamma_dev, const float* beta_dev,
                      float eps, int silu, const float* dy_dev, const float* add_dev, float* dx_f32_dev,
                      void* dx_f16_dev, void* stream);
int dfb_layernorm_bwd(const float* x_dev, int rows, int C, const float* gamma_dev, float eps, const float* dy_dev,
                      const float* add_dev, float* dx_f32_dev, void* dx_f16_dev, void* stream);
/* softmax(q k^T scale) v backward per (sample, head), d <= 64: q/k/v/o fp16 (head h at columns h*d), dO fp32;
 * dq/dk/dv fp16 (dk, dv may be NULL: cross-attention against a constant context); lse_ws / d_ws: fp32
 * [B*heads*Lq] scratch.  Deterministic (no atomics). */
int dfb_attention_bwd(const void* q_dev, int ldq, const void* k_dev, int ldk, const void* v_dev, int ldv,
                      const void* o_dev, int ldo, const float* dO_dev, int lddo, int B, int heads, int Lq, int Lk,
                      int d, float scale, void* dq_dev, int lddq, void* dk_dev, int lddk, void* dv_dev, int lddv,
                      float* lse_ws_dev, float* d_ws_dev, void* stream);
/* GEGLU on an un-interleaved projection fp32 [M, 2F] = [value | gate] (attention_openai.py:37-44) */
int dfb_geglu_fwd(const float* proj_dev, long long M, int F, void* h_f16_dev, void* stream);
int dfb_geglu_bwd(const float* proj_dev, const float* dh_dev, long long M, int F, void* dproj_f16_dev, void* stream);
/* backward-data scatter of the 3x3 / stride-2 Downsample conv as a gather: dcol fp32 [B*(H/2)*(W/2), 9*C] */
int dfb_col2im_s2(const float* dcol_dev, int B, int H, int W, int C, const float* add_dev, float* dx_f32_dev,
                  void* dx_f16_dev, void* stream);
/* avg-pool + Linear(C -> 1) + sigmoid (alignment_backbone.py:676-686) and the seed of d log(prob) / dc * seed_scale */
int dfb_classifier_head(const float* c_dev, int B, int HW, int C, const float* w_dev, const float* bias_dev,
                        float seed_scale, float* prob_dev, void* dc_f16_dev, void* stream);
int dfb_scale_f32(float* x_dev, float s, long long n, void* stream);
int dfb_cast_f16(const float* src_dev, void* dst_f16_dev, long long n, void* stream);
/* the 4-channel boundary convs: stem (NCHW fp32 -> NHWC fp32, w packed [(ci*9+tap), Cout]) and head (NHWC fp16 ->
 * NCHW fp32, Cout <= 4, w packed [Cout, 9, C]); 3x3, pad 1 */
int dfb_stem_conv(const float* x_nchw_dev, int B, int Cin, int H, int W, const float* w_packed_dev,
                  const float* bias_dev, int Cout, float* out_nhwc_dev, void* stream);
int dfb_head_conv(const void* a_f16_nhwc_dev, int B, int H, int W, int C, const float* w_packed_dev,
                  const float* bias_dev, int Cout, float* out_nchw_dev, void* stream);
/* CAVP frame ingest (reference inference/demo_util.py:135-163: cv2 BGR->RGB, PIL Resize((224,224)), ToTensor).
 * src: N uint8 frames [N,H,W,3] on the device (swap_rb = 1 for cv2's BGR order).  kk_* / bounds_* are the
 * int32 coefficient tables of Pillow's resample for (W -> OW) and (H -> OH) (precompute_coeffs +
 * normalize_coeffs_8bpc, built by diff_foley_b200/frames.py), on the device: kk [O, ksize], bounds [O, 2] =
 * (first source index, count).  tmp_u8: [N,H,OW,3] scratch.  out_f32: [N,3,OH,OW] in [0,1] -- bit-identical
 * to ToTensor()(PIL resize); out_u8 (optional): the resized uint8 image [N,OH,OW,3]. */
int dfb_frames_resize(const void* src_u8_dev, int N, int H, int W, int swap_rb, const int32_t* kk_h_dev,
                      const int32_t* bounds_h_dev, int ksize_h, int OW, const int32_t* kk_v_dev,
                      const int32_t* bounds_v_dev, int ksize_v, int OH, void* tmp_u8_dev, float* out_f32_dev,
                      void* out_u8_dev, void* stream);
/* Multi-GPU (one process per GPU): the library owns its NCCL communicator, created from a 128-byte
 * ncclUniqueId that rank 0 obtains with dfb_comm_unique_id and the host distributes (e.g. a
 * torch.distributed broadcast); freed by dfb_comm_destroy / dfb_unet_destroy. */
int dfb_comm_unique_id(void* id128_out);
int dfb_comm_init(dfb_handle h, int rank, int world, const void* id128);
int dfb_comm_destroy(dfb_handle h);
/* Per-launch profile of one UNet forward: runs the plan for b_eff `iters` times with a CUDA event
 * between consecutive launches (on `stream`) and returns, per launch, the kernel kind, its
 * algorithmic FLOPs / bytes and the mean event-timed duration.  Context must have been set. */
typedef struct dfb_op_info {
  char kind[24];
  int32_t M, N, K, splits, ctas;
  double flops;
  double bytes;
  float ms;
} dfb_op_info;
int dfb_unet_profile(dfb_handle h, const float* x_dev, int x_repeat, const void* t_dev, int t_is_float,
                     float* out_dev, int b_eff, int iters, dfb_op_info* infos, int cap, int* n_ops,
                     void* stream);
/* Tuning aid (tools/autotune_igemm.py): force the GEMM tile width (64 / 128; 0 = planner's choice) and
 * operand-ring depth (1 deep / 0 shallow; -1 = default) of every GEMM planned afterwards. */
void dfb_debug_igemm_force(int bn, int deep);
/* CTA-pair tiles (tcgen05.mma.cta_group::2, 256 x 128 per SM pair) for every GEMM planned afterwards that can take
 * them (>= 2 M tiles, N >= 128, no split-K): 1 on, 0 off, -1 = follow the DFB_PAIR environment variable (default off:
 * measured +3-5 % on the large-K convolutions at B_eff = 16, -5-15 % on the short-K Linears) */
void dfb_debug_igemm_pair(int pair);
/* In-kernel timeline of one graph-replayed UNet forward (diagnostics): for launch i and mark k,
 * marks[32*i + 2*k] is the first and ~marks[32*i + 2*k + 1] the last %globaltimer reading (ns) at
 * which a CTA of that launch passed the mark; mark 0 = kernel entry, 1 = programmatic-dependent-
 * launch wait released, 2..6 = GEMM phases (first operand stage landed, last MMA issued, accumulator
 * complete, epilogue done, split-K reduction done), 7 = exit, 8..15 = finer epilogue marks.  All-ones where a mark was not hit.
 * `marks` is a HOST array of 32*cap words.  Context must have been set. */
int dfb_unet_trace(dfb_handle h, const float* x_dev, int x_repeat, const void* t_dev, int t_is_float,
                   float* out_dev, int b_eff, unsigned long long* marks, int cap, int* n_ops,
                   void* stream);
/* Debug/test aid: block outputs ("input_blocks.3", "middle_block", ...) of the plan for b_eff,
 * channels-last fp32 [b_eff, H, W, C].  With the environment variable DFB_DEBUG_TAPS=1 set before the
 * first forward every block output keeps its own buffer; otherwise only the skip-stack tensors
 * (input blocks) are still intact after a forward. */
int dfb_unet_debug_num_taps(dfb_handle h, int b_eff);
int dfb_unet_debug_tap(dfb_handle h, int b_eff, int i, char* name, int name_cap, int32_t* hwc,
                       float* dst_dev, void* stream);
/* kernels launched by the most recent forward / sample call (for bench.py's gpu_launches) */
long long dfb_unet_last_launch_count(dfb_handle h);
int dfb_unet_destroy(dfb_handle h);

/* --------------------------------------------------------------- per-kernel entry points
 * Used by the parity tests (tests/test_ops_gpu.py) and by callers that want single ops. */

/* out[M,N] = act(A[M,K] . W[N,K]^T + bias + residual); A, W fp16 row-major; tcgen05 tensor cores.
 * splits: 0 = auto split-K, >=1 forced.  (nn.Linear / 1x1 Conv2d) */
int dfb_gemm(const void* a_f16_dev, const void* w_f16_dev, int M, int N, int K, const float* bias_dev,
             const float* residual_dev, int act, float* out_f32_dev, void* out_f16_dev, int splits,
             void* stream);
/* LayerNorm folded into the consuming Linear (attention_openai.py:211-215: attn(norm(x)), ff(norm(x))).
 * Producer side: dfb_gemm_stats is dfb_gemm that also writes, per output row and per N-tile of the launch,
 * the partial (sum, sum of squares) of its fp32 outputs to stats_dev [M, *tiles_n_out] (float2); tiles_n_out
 * is a host int the call fills in.  Consumer side: dfb_gemm_ln computes
 *   out = act( rstd_m * (A . W'^T) - rstd_m * mu_m * ln_s[n] + t[n] )
 * with A the raw fp16 x, W' = fp16(gamma . W), ln_s[n] = sum_k W'[n,k], t[n] = sum_k beta_k W[n,k] + b[n],
 * and (mu_m, rstd_m) merged from ln_stats_dev [M, ln_tiles] over C = K columns with eps ln_eps. */
int dfb_gemm_stats(const void* a_f16_dev, const void* w_f16_dev, int M, int N, int K, const float* bias_dev,
                   const float* residual_dev, float* out_f32_dev, void* out_f16_dev, int splits,
                   void* stats_dev, int* tiles_n_out, void* stream);
int dfb_gemm_ln(const void* a_f16_dev, const void* w_f16_dev, int M, int N, int K, const float* t_dev,
                const float* ln_s_dev, const void* ln_stats_dev, int ln_tiles, float ln_eps, int act,
                float* out_f32_dev, void* out_f16_dev, int splits, void* stream);
/* 3x3 / stride 1 / pad 1 convolution as implicit GEMM; a: fp16 NHWC [B,H,W,C], w: fp16 [N, 9*C]
 * with k = (ky*3+kx)*C + c; rowvec: optional fp32 [B,N] added per sample (timestep embedding,
 * openai_unetmodel.py:263-272); residual: optional fp32 NHWC [B,H,W,N]. */
int dfb_conv3x3(const void* a_f16_dev, const void* w_f16_dev, int B, int H, int W, int C, int N,
                const float* bias_dev, const float* rowvec_dev, const float* residual_dev, int act,
                float* out_f32_dev, void* out_f16_dev, int splits, void* stream);
/* 3x3 conv (pad 1) over fp16 NHWC `a` [B,H,W,C] plus a 1x1 conv over a second fp16 NHWC tensor `a2`
 * [B,H,W,C2] accumulated into the same output: weights fp16 [N, 9*C + C2] (the 1x1 weights are the last C2
 * columns), biases `bias` + `bias2` (either may be NULL).  This is ResBlock.out_layers conv + skip_connection
 * (openai_unetmodel.py:236-243,275) as ONE implicit GEMM. */
int dfb_conv3x3_cat(const void* a, const void* a2, int C2, const void* w, int B, int H, int W, int C, int N,
                    const float* bias, const float* bias2, const float* residual, float* out_f32,
                    void* out_f16, int splits, void* stream);
/* "same"-padded stride-1 convolution with an odd (kt,kh,kw) kernel, kt*kh*kw <= 9, over fp16
 * channels-last [B,T,H,W,C] as implicit GEMM (one TMA box per tap); w: fp16 [N, taps*C] with
 * k = ((it*kh + ih)*kw + iw)*C + c.  Covers the CAVP encoders' (1,1,1), (1,3,3), (3,1,1) Conv3d and 3x3
 * Conv2d with BatchNorm folded into w / bias (inference/model/cavp_modules.py:243-293, 1440-1483);
 * residual_f16: optional fp16 [B,T,H,W,N] identity path, added before the activation. */
int dfb_conv_taps(const void* a_f16_dev, const void* w_f16_dev, int B, int T, int H, int W, int C, int N,
                  int kt, int kh, int kw, const float* bias_dev, const void* residual_f16_dev, int act,
                  float* out_f32_dev, void* out_f16_dev, int splits, void* stream);
/* im2col on fp16 channels-last images [NI,H,W,C] -> [NI*Ho*Wo, Kpad] (k = (ky*kw+kx)*C + c, zero padded
 * to Kpad): strided / large-kernel convs (7x7 stem, stride-2 3x3 and 1x1) become plain GEMMs. */
int dfb_im2col_f16(const void* src_dev, void* dst_dev, int NI, int H, int W, int C, int kh, int kw,
                   int stride, int pad, int Kpad, void* stream);
/* max (is_max=1) or average pooling on fp16 channels-last images */
int dfb_pool2d_f16(const void* src_dev, void* dst_dev, int NI, int H, int W, int C, int kh, int kw, int sh,
                   int sw, int ph, int pw, int is_max, void* stream);
/* GroupNorm(32) (+SiLU) over concat(src0, src1) channels-last fp32 -> fp16 (util.py:214-216); group slabs
 * beyond 64 K elements (the first-stage decoder's upper levels) take a two-kernel statistics + apply path */
int dfb_groupnorm(const float* src0_dev, int C0, const float* src1_dev, int C1, int B, int HW,
                  const float* gamma_dev, const float* beta_dev, float eps, int silu, void* out_f16_dev,
                  void* raw_f16_dev, void* stream);
/* softmax(scale * x) over the rows of fp32 [rows, n] (n % 4 == 0, n <= 2048) -> fp16: the first-stage
 * decoder's 512-wide single-head attention (stage1_autoencoder/model.py:245-300) between two dfb_gemm calls */
int dfb_softmax_rows(const float* src_dev, int rows, int n, float scale, void* out_f16_dev, void* stream);
int dfb_layernorm(const float* src_dev, int rows, int C, const float* gamma_dev, const float* beta_dev,
                  float eps, void* out_f16_dev, void* stream);
/* softmax(q k^T scale) v per (sample, head) (attention_openai.py:170-193); fp16 in/out */
int dfb_attention(const void* q_dev, int ldq, const void* k_dev, int ldk, const void* v_dev, int ldv,
                  void* out_dev, int ldo, int B, int heads, int Lq, int Lk, int d, int dpad, float scale,
                  void* stream);
/* sinusoidal timestep embedding (util.py:151-171) -> fp16 [B, dim] */
int dfb_temb(const void* t_dev, int t_is_float, int B, int dim, void* out_f16_dev, void* stream);
int dfb_upsample2x_f16(const float* src_dev, void* dst_f16_dev, int B, int H, int W, int C, void* stream);
int dfb_im2col_s2(const float* src_dev, void* dst_f16_dev, int B, int H, int W, int C, void* stream);
/* one DDIM update with CFG combine (ddim.py:241-245, 258-273); eps_uncond/grad may be NULL */
int dfb_ddim_step(const float* x_dev, const float* eps_uncond_dev, const float* eps_cond_dev,
                  const float* grad_dev, float cfg_scale, float sqrt_one_minus_at, float sqrt_at,
                  float sqrt_a_prev, float dir_coef, float grad_coef, float* x_prev_dev,
                  float* pred_x0_dev, size_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DFB_H_ */
