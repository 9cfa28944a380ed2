// capi.cu -- per-kernel C entry points (include/dfb.h, "per-kernel entry points").  They wrap the
// same launchers the UNet engine uses, so the parity tests exercise exactly the shipped kernels.
#include <cstdlib>
#include <mutex>

#include "../../include/dfb.h"
#include "dfb_internal.h"

namespace dfb {

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DFB_NO_PDL");
    v = (e && e[0] == '1') ? 0 : 1;
  }
  return v == 1;
}

int kernels_init() {
  static std::once_flag once;
  static int rc = 0;
  std::call_once(once, [] {
    rc = igemm_init();
    if (!rc) rc = norm_init();
    if (!rc) rc = attention_init();
    if (!rc) rc = elementwise_init();
  });
  return rc;
}

static int run_igemm(const __half* a, const __half* w, int N, const IGemmGeom& g, IGemmEpilogue ep,
                     int splits, cudaStream_t s) {
  int r = kernels_init();
  if (r) return r;
  IGemmPlan plan;
  r = igemm_plan(&plan, a, w, N, g, ep, splits);
  if (r) return r;
  return igemm_launch(plan, s);
}

}  // namespace dfb

using namespace dfb;

extern "C" {

int dfb_gemm(const void* a, const void* w, int M, int N, int K, const float* bias, const float* residual,
             int act, float* out_f32, void* out_f16, int splits, void* stream) {
  if (!a || !w || M < 1 || N < 1 || K < 1) { set_error("dfb_gemm: bad argument"); return DFB_E_INVALID; }
  IGemmEpilogue ep;
  memset(&ep, 0, sizeof(ep));
  ep.out_f32 = out_f32;
  ep.out_f16 = (__half*)out_f16;
  ep.ldo = (act == ACT_GEGLU) ? N / 2 : N;
  ep.bias = bias;
  ep.residual = residual;
  ep.ld_res = ep.ldo;
  ep.act = act;
  return run_igemm((const __half*)a, (const __half*)w, N, gemm_geom(M, K), ep, splits, (cudaStream_t)stream);
}

int dfb_gemm_stats(const void* a, const void* w, int M, int N, int K, const float* bias, const float* residual,
                   float* out_f32, void* out_f16, int splits, void* stats, int* tiles_n_out, void* stream) {
  if (!a || !w || !stats || !tiles_n_out || M < 1 || N < 1 || K < 1) { set_error("dfb_gemm_stats: bad argument"); return DFB_E_INVALID; }
  IGemmEpilogue ep;
  memset(&ep, 0, sizeof(ep));
  ep.out_f32 = out_f32;
  ep.out_f16 = (__half*)out_f16;
  ep.ldo = N;
  ep.bias = bias;
  ep.residual = residual;
  ep.ld_res = N;
  ep.stats_out = (float2*)stats;
  int r = kernels_init();
  if (r) return r;
  IGemmPlan plan;
  r = igemm_plan(&plan, (const __half*)a, (const __half*)w, N, gemm_geom(M, K), ep, splits);
  if (r) return r;
  *tiles_n_out = plan.tiles_n;
  return igemm_launch(plan, (cudaStream_t)stream);
}

int dfb_gemm_ln(const void* a, const void* w, int M, int N, int K, const float* t, const float* ln_s,
                const void* ln_stats, int ln_tiles, float ln_eps, int act, float* out_f32, void* out_f16,
                int splits, void* stream) {
  if (!a || !w || !t || !ln_s || !ln_stats || M < 1 || N < 1 || K < 1 || ln_tiles < 1) { set_error("dfb_gemm_ln: bad argument"); return DFB_E_INVALID; }
  IGemmEpilogue ep;
  memset(&ep, 0, sizeof(ep));
  ep.out_f32 = out_f32;
  ep.out_f16 = (__half*)out_f16;
  ep.ldo = (act == ACT_GEGLU) ? N / 2 : N;
  ep.bias = t;
  ep.act = act;
  ep.ln_stats = (const float2*)ln_stats;
  ep.ln_tiles = ln_tiles;
  ep.ln_inv_c = 1.0f / (float)K;
  ep.ln_eps = ln_eps;
  ep.ln_s = ln_s;
  return run_igemm((const __half*)a, (const __half*)w, N, gemm_geom(M, K), ep, splits, (cudaStream_t)stream);
}

int dfb_conv3x3(const void* a, const void* w, int B, int H, int W, int C, int N, const float* bias,
                const float* rowvec, const float* residual, int act, float* out_f32, void* out_f16,
                int splits, void* stream) {
  if (!a || !w || B < 1 || H < 1 || W < 1) { set_error("dfb_conv3x3: bad argument"); return DFB_E_INVALID; }
  IGemmEpilogue ep;
  memset(&ep, 0, sizeof(ep));
  ep.out_f32 = out_f32;
  ep.out_f16 = (__half*)out_f16;
  ep.ldo = N;
  ep.bias = bias;
  ep.rowvec = rowvec;
  ep.ld_rowvec = N;
  ep.rows_per_sample = H * W;
  ep.residual = residual;
  ep.ld_res = N;
  ep.act = act;
  return run_igemm((const __half*)a, (const __half*)w, N, conv3x3_geom(B, H, W, C), ep, splits,
                   (cudaStream_t)stream);
}

int dfb_conv3x3_cat(const void* a, const void* a2, int C2, const void* w, int B, int H, int W, int C, int N,
                    const float* bias, const float* bias2, const float* residual, float* out_f32,
                    void* out_f16, int splits, void* stream) {
  if (!a || !a2 || !w || B < 1 || H < 1 || W < 1 || C2 < 1) { set_error("dfb_conv3x3_cat: bad argument"); return DFB_E_INVALID; }
  IGemmEpilogue ep;
  memset(&ep, 0, sizeof(ep));
  ep.out_f32 = out_f32;
  ep.out_f16 = (__half*)out_f16;
  ep.ldo = N;
  ep.bias = bias;
  ep.rowvec = bias2;  // a second bias vector: "per-sample vector" with stride 0
  ep.ld_rowvec = 0;
  ep.rows_per_sample = H * W;
  ep.residual = residual;
  ep.ld_res = N;
  IGemmGeom g = conv3x3_geom(B, H, W, C);
  g.C2 = C2;
  g.A2 = (const __half*)a2;
  return run_igemm((const __half*)a, (const __half*)w, N, g, ep, splits, (cudaStream_t)stream);
}

int dfb_conv_taps(const void* a, const void* w, int B, int T, int H, int W, int C, int N, int kt, int kh,
                  int kw, const float* bias, const void* residual_f16, int act, float* out_f32,
                  void* out_f16, int splits, void* stream) {
  if (!a || !w || kt * kh * kw > 9 || kt < 1 || kh < 1 || kw < 1 || !(kt & 1) || !(kh & 1) || !(kw & 1)) {
    set_error("dfb_conv_taps: odd kernel extents with kt*kh*kw <= 9 required");
    return DFB_E_INVALID;
  }
  IGemmEpilogue ep;
  memset(&ep, 0, sizeof(ep));
  ep.out_f32 = out_f32;
  ep.out_f16 = (__half*)out_f16;
  ep.ldo = N;
  ep.bias = bias;
  ep.residual_f16 = (const __half*)residual_f16;
  ep.ld_res = N;
  ep.act = act;
  return run_igemm((const __half*)a, (const __half*)w, N, conv_taps_geom(B, T, H, W, C, kt, kh, kw), ep,
                   splits, (cudaStream_t)stream);
}

void dfb_debug_igemm_force(int bn, int deep) { igemm_force(bn, deep); }
void dfb_debug_igemm_pair(int pair) { igemm_force_pair(pair); }

// ---- backward / classifier ops (backward.cu) and the small-channel boundary convs
int dfb_groupnorm_bwd(const float* x, int C, int B, int HW, const float* gamma, const float* beta, float eps, int silu,
                      const float* dy, const float* add, float* dx_f32, void* dx_f16, void* stream) {
  if (!x || !gamma || !beta || !dy || (!dx_f32 && !dx_f16)) { set_error("dfb_groupnorm_bwd: null argument"); return DFB_E_INVALID; }
  int r = kernels_init();
  if (r) return r;
  return groupnorm_bwd_launch(x, C, B, HW, gamma, beta, eps, silu, dy, add, dx_f32, (__half*)dx_f16, (cudaStream_t)stream);
}
int dfb_layernorm_bwd(const float* x, int rows, int C, const float* gamma, float eps, const float* dy, const float* add,
                      float* dx_f32, void* dx_f16, void* stream) {
  if (!x || !gamma || !dy || (!dx_f32 && !dx_f16)) { set_error("dfb_layernorm_bwd: null argument"); return DFB_E_INVALID; }
  int r = kernels_init();
  if (r) return r;
  return layernorm_bwd_launch(x, rows, C, gamma, eps, dy, add, dx_f32, (__half*)dx_f16, (cudaStream_t)stream);
}
int dfb_attention_bwd(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, const void* o, int ldo,
                      const float* dO, int lddo, int B, int heads, int Lq, int Lk, int d, float scale, void* dq, int lddq,
                      void* dk, int lddk, void* dv, int lddv, float* lse_ws, float* d_ws, void* stream) {
  if (!q || !k || !v || !o || !dO || !dq || !lse_ws || !d_ws) { set_error("dfb_attention_bwd: null argument"); return DFB_E_INVALID; }
  int r = kernels_init();
  if (r) return r;
  return attention_bwd_launch((const __half*)q, ldq, (const __half*)k, ldk, (const __half*)v, ldv, (const __half*)o, ldo,
                              dO, lddo, B, heads, Lq, Lk, d, scale, (__half*)dq, lddq, (__half*)dk, lddk, (__half*)dv,
                              lddv, lse_ws, d_ws, (cudaStream_t)stream);
}
int dfb_geglu_fwd(const float* proj, long long M, int F, void* h_f16, void* stream) {
  if (!proj || !h_f16) { set_error("dfb_geglu_fwd: null argument"); return DFB_E_INVALID; }
  int r = kernels_init();
  if (r) return r;
  return geglu_fwd_launch(proj, (long)M, F, (__half*)h_f16, (cudaStream_t)stream);
}
int dfb_geglu_bwd(const float* proj, const float* dh, long long M, int F, void* dproj_f16, void* stream) {
  if (!proj || !dh || !dproj_f16) { set_error("dfb_geglu_bwd: null argument"); return DFB_E_INVALID; }
  int r = kernels_init();
  if (r) return r;
  return geglu_bwd_launch(proj, dh, (long)M, F, (__half*)dproj_f16, (cudaStream_t)stream);
}
int dfb_col2im_s2(const float* dcol, int B, int H, int W, int C, const float* add, float* dx_f32, void* dx_f16, void* stream) {
  if (!dcol || (!dx_f32 && !dx_f16) || (H & 1) || (W & 1)) { set_error("dfb_col2im_s2: bad argument"); return DFB_E_INVALID; }
  int r = kernels_init();
  if (r) return r;
  return col2im_s2_launch(dcol, B, H, W, C, add, dx_f32, (__half*)dx_f16, (cudaStream_t)stream);
}
int dfb_classifier_head(const float* c, int B, int HW, int C, const float* w, const float* bias, float seed_scale,
                        float* prob, void* dc_f16, void* stream) {
  if (!c || !w || !bias) { set_error("dfb_classifier_head: null argument"); return DFB_E_INVALID; }
  int r = kernels_init();
  if (r) return r;
  return classifier_head_launch(c, B, HW, C, w, bias, seed_scale, prob, (__half*)dc_f16, (cudaStream_t)stream);
}
int dfb_scale_f32(float* x, float s, long long n, void* stream) {
  if (!x) { set_error("dfb_scale_f32: null argument"); return DFB_E_INVALID; }
  int r = kernels_init();
  if (r) return r;
  return scale_f32_launch(x, s, (long)n, (cudaStream_t)stream);
}
int dfb_cast_f16(const float* src, void* dst_f16, long long n, void* stream) {
  if (!src || !dst_f16) { set_error("dfb_cast_f16: null argument"); return DFB_E_INVALID; }
  int r = kernels_init();
  if (r) return r;
  return cast_f16_launch(src, (__half*)dst_f16, (size_t)n, (cudaStream_t)stream);
}
int dfb_stem_conv(const float* x_nchw, int B, int Cin, int H, int W, const float* w_packed, const float* bias, int Cout,
                  float* out_nhwc, void* stream) {
  if (!x_nchw || !w_packed || !bias || !out_nhwc) { set_error("dfb_stem_conv: null argument"); return DFB_E_INVALID; }
  int r = kernels_init();
  if (r) return r;
  return stem_conv_launch(x_nchw, B, 0, B, Cin, H, W, w_packed, bias, Cout, out_nhwc, (cudaStream_t)stream);
}
int dfb_head_conv(const void* a_f16_nhwc, int B, int H, int W, int C, const float* w_packed, const float* bias, int Cout,
                  float* out_nchw, void* stream) {
  if (!a_f16_nhwc || !w_packed || !bias || !out_nchw || Cout > 4) { set_error("dfb_head_conv: bad argument"); return DFB_E_INVALID; }
  int r = kernels_init();
  if (r) return r;
  return head_conv_launch((const __half*)a_f16_nhwc, B, H, W, C, w_packed, bias, Cout, out_nchw, (cudaStream_t)stream);
}

int dfb_frames_resize(const void* src_u8, int N, int H, int W, int swap_rb, const int32_t* kk_h, const int32_t* bounds_h,
                      int ksize_h, int OW, const int32_t* kk_v, const int32_t* bounds_v, int ksize_v, int OH,
                      void* tmp_u8, float* out_f32, void* out_u8, void* stream) {
  if (!src_u8 || !kk_h || !bounds_h || !kk_v || !bounds_v || !tmp_u8 || !out_f32 || N < 1 || H < 1 || W < 1 || OW < 1 ||
      OH < 1 || ksize_h < 1 || ksize_v < 1) { set_error("dfb_frames_resize: bad argument"); return DFB_E_INVALID; }
  int r = kernels_init();
  if (r) return r;
  return frames_resize_launch((const uint8_t*)src_u8, N, H, W, swap_rb, kk_h, bounds_h, ksize_h, OW, kk_v, bounds_v,
                              ksize_v, OH, (uint8_t*)tmp_u8, out_f32, (uint8_t*)out_u8, (cudaStream_t)stream);
}

int dfb_im2col_f16(const void* src, void* dst, int NI, int H, int W, int C, int kh, int kw, int stride,
                   int pad, int Kpad, void* stream) {
  return im2col_f16_launch((const __half*)src, (__half*)dst, NI, H, W, C, kh, kw, stride, pad, Kpad,
                           (cudaStream_t)stream);
}

int dfb_pool2d_f16(const void* src, void* dst, int NI, int H, int W, int C, int kh, int kw, int sh, int sw,
                   int ph, int pw, int is_max, void* stream) {
  return pool2d_f16_launch((const __half*)src, (__half*)dst, NI, H, W, C, kh, kw, sh, sw, ph, pw, is_max,
                           (cudaStream_t)stream);
}

int dfb_groupnorm(const float* src0, int C0, const float* src1, int C1, int B, int HW, const float* gamma,
                  const float* beta, float eps, int silu, void* out, void* raw, void* stream) {
  int r = kernels_init();
  if (r) return r;
  return groupnorm_launch(src0, C0, src1, C1, B, HW, gamma, beta, eps, silu, (__half*)out, (__half*)raw,
                          (cudaStream_t)stream);
}

int dfb_layernorm(const float* src, int rows, int C, const float* gamma, const float* beta, float eps,
                  void* out, void* stream) {
  return layernorm_launch(src, rows, C, gamma, beta, eps, (__half*)out, (cudaStream_t)stream);
}

int dfb_softmax_rows(const float* src, int rows, int n, float scale, void* out, void* stream) {
  return softmax_rows_launch(src, rows, n, scale, (__half*)out, (cudaStream_t)stream);
}

int dfb_attention(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out,
                  int ldo, int B, int heads, int Lq, int Lk, int d, int dpad, float scale, void* stream) {
  int r = kernels_init();
  if (r) return r;
  return attention_launch((const __half*)q, ldq, (const __half*)k, ldk, (const __half*)v, ldv,
                          (__half*)out, ldo, B, heads, Lq, Lk, d, dpad, scale, (cudaStream_t)stream);
}

int dfb_temb(const void* t, int t_is_float, int B, int dim, void* out, void* stream) {
  return temb_launch(t, t_is_float, B, dim, (__half*)out, (cudaStream_t)stream);
}

int dfb_upsample2x_f16(const float* src, void* dst, int B, int H, int W, int C, void* stream) {
  return upsample2x_f16_launch(src, (__half*)dst, B, H, W, C, (cudaStream_t)stream);
}

int dfb_im2col_s2(const float* src, void* dst, int B, int H, int W, int C, void* stream) {
  return im2col_s2_launch(src, (__half*)dst, B, H, W, C, (cudaStream_t)stream);
}

int dfb_ddim_step(const float* x, const float* eu, const float* ec, const float* grad, float cfg_scale,
                  float sqrt_one_minus_at, float sqrt_at, float sqrt_a_prev, float dir_coef,
                  float grad_coef, float* x_prev, float* pred_x0, size_t n, void* stream) {
  if (!x || !ec || !x_prev) { set_error("dfb_ddim_step: null argument"); return DFB_E_INVALID; }
  return ddim_update_launch(x, eu, ec, grad, cfg_scale, sqrt_one_minus_at, sqrt_at, sqrt_a_prev, dir_coef,
                            grad_coef, x_prev, pred_x0, n, (cudaStream_t)stream);
}

}  // extern "C"
