"""Per-kernel parity: each CUDA kernel, called through the C ABI, against plain torch fp32 on the
same (fp16-rounded where the kernel consumes fp16) inputs.  Tolerances are written per test.

rel-L2 = ||y - y_ref|| / ||y_ref||.
"""
import math

import pytest
import torch
import torch.nn.functional as F

from diff_foley_b200 import _lib as L

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def sync():
    torch.cuda.synchronize()


def gemm(a16, w16, bias=None, residual=None, act=0, out_dtype=torch.float32, splits=0):
    M, K = a16.shape
    N = w16.shape[0]
    No = N // 2 if act == 2 else N
    out = torch.full((M, No), float("nan"), device=DEV, dtype=out_dtype)
    f32 = out if out_dtype == torch.float32 else None
    f16 = out if out_dtype == torch.float16 else None
    L.check(L.lib().dfb_gemm(L.ptr(a16), L.ptr(w16), M, N, K, L.ptr(bias), L.ptr(residual), act,
                             L.ptr(f32), L.ptr(f16), splits, L.cur_stream()), "dfb_gemm")
    sync()
    return out


# shapes cover: tiny M (time-embed), M not a multiple of 128, N=320 (BN=64 path), K with many
# k-blocks, the biggest transformer GEMMs, forced split-K
@pytest.mark.parametrize("M,N,K,splits", [
    (2, 1280, 320, 0), (128, 128, 64, 1), (200, 320, 320, 1), (2048, 320, 320, 0),
    (512, 640, 2560, 0), (128, 1280, 5120, 0), (32, 1280, 1280, 0), (2048, 960, 320, 1),
    (256, 256, 1024, 4), (64, 25600, 768, 1),
])
def test_gemm_plain(M, N, K, splits):
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K)
    a = torch.randn(M, K, generator=g).to(DEV).half()
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(DEV).half()
    ref = a.float() @ w.float().t()
    out = gemm(a, w, splits=splits)
    assert torch.isfinite(out).all()
    # fp16 operands are exact in both; only fp32 accumulation order differs
    assert rel_l2(out, ref) < 2e-6, (M, N, K)


def test_gemm_epilogues():
    g = torch.Generator(device="cpu").manual_seed(11)
    M, N, K = 384, 640, 640
    a = torch.randn(M, K, generator=g).to(DEV).half()
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(DEV).half()
    bias = torch.randn(N, generator=g).to(DEV)
    res = torch.randn(M, N, generator=g).to(DEV)
    base = a.float() @ w.float().t() + bias
    out = gemm(a, w, bias=bias, residual=res)
    assert rel_l2(out, base + res) < 2e-6
    out = gemm(a, w, bias=bias, act=1, out_dtype=torch.float16)
    assert rel_l2(out.float(), F.silu(base)) < 1e-3  # fp16 output rounding
    out = gemm(a, w, bias=bias, residual=res, splits=3)
    assert rel_l2(out, base + res) < 2e-6


@pytest.mark.parametrize("splits", [1, 2])
def test_gemm_geglu(splits):
    """GEGLU epilogue (attention_openai.py:37-44) with the per-128-column value|gate interleave."""
    g = torch.Generator(device="cpu").manual_seed(5)
    M, C = 256, 320
    a = torch.randn(M, C, generator=g).to(DEV).half()
    w = (torch.randn(8 * C, C, generator=g) / math.sqrt(C)).to(DEV).half()
    bias = torch.randn(8 * C, generator=g).to(DEV)
    proj = a.float() @ w.float().t() + bias
    x, gate = proj.chunk(2, dim=-1)
    ref = x * F.gelu(gate)
    # interleave rows: tile t of 128 = 64 value rows | 64 gate rows
    C4 = 4 * C
    wv, wg = w[:C4].view(C4 // 64, 64, C), w[C4:].view(C4 // 64, 64, C)
    wi = torch.cat([wv, wg], dim=1).reshape(8 * C, C).contiguous()
    bv, bg = bias[:C4].view(-1, 64), bias[C4:].view(-1, 64)
    bi = torch.cat([bv, bg], dim=1).reshape(-1).contiguous()
    out = gemm(a, wi, bias=bi, act=2, out_dtype=torch.float16, splits=splits)
    assert out.shape == (M, C4)
    assert rel_l2(out.float(), ref) < 1e-3


# LayerNorm folded into the consuming GEMM: producer (x = A.Wp^T + b + res, fp32 + fp16 copy + per-tile row
# statistics) followed by consumer (LN(x).W^T + b as rstd*(x16.W'^T) - rstd*mu*s + t); the UNet's
# (tokens, C) call sites incl. the M = 32 level (3/4 of the tile rows outside the tensor), split-K both ways
@pytest.mark.parametrize("M,C,N,act,sp_p,sp_c", [
    (2048, 320, 1152, 0, 0, 0), (512, 640, 1920, 0, 0, 0), (128, 1280, 3840, 0, 0, 0), (32, 1280, 3840, 0, 0, 0),
    (32, 1280, 1280, 0, 5, 4), (2048, 320, 2560, 2, 1, 1), (128, 1280, 10240, 2, 0, 0), (32, 1280, 10240, 2, 0, 0),
    (200, 64, 192, 0, 1, 2),
])
def test_gemm_layernorm_fold(M, C, N, act, sp_p, sp_c):
    import ctypes as Ct
    g = torch.Generator(device="cpu").manual_seed(M + C + N + act)
    a = torch.randn(M, C, generator=g).to(DEV).half()
    wp = (torch.randn(C, C, generator=g) / math.sqrt(C)).to(DEV).half()
    bp = torch.randn(C, generator=g).to(DEV)
    res = (torch.randn(M, C, generator=g) + 0.5).to(DEV)
    gamma = (1.0 + 0.3 * torch.randn(C, generator=g)).to(DEV)
    beta = (0.2 * torch.randn(C, generator=g)).to(DEV)
    w = (torch.randn(N, C, generator=g) / math.sqrt(C)).to(DEV)
    b = torch.randn(N, generator=g).to(DEV)
    lib = L.lib()
    # ---- producer
    x32 = torch.full((M, C), float("nan"), device=DEV)
    x16 = torch.zeros(M, C, device=DEV, dtype=torch.float16)
    stats = torch.zeros(M, 64, 2, device=DEV)
    tiles = Ct.c_int(0)
    L.check(lib.dfb_gemm_stats(L.ptr(a), L.ptr(wp), M, C, C, L.ptr(bp), L.ptr(res), L.ptr(x32), L.ptr(x16), sp_p,
                               L.ptr(stats), Ct.byref(tiles), L.cur_stream()), "dfb_gemm_stats")
    sync()
    x_ref = a.float() @ wp.float().t() + bp + res
    assert rel_l2(x32, x_ref) < 2e-6
    st = stats.view(-1)[: M * tiles.value * 2].view(M, tiles.value, 2)
    assert rel_l2(st[..., 0].sum(1), x32.sum(1)) < 1e-5 and rel_l2(st[..., 1].sum(1), (x32 * x32).sum(1)) < 1e-5
    # ---- consumer: folded weights as engine.cu's ln_fold_kernel builds them
    wf = (gamma[None, :] * w).half()
    s_n = wf.float().sum(1)
    t_n = (w * beta[None, :]).sum(1) + b
    if act == 2:   # GEGLU: value | gate interleaved per 128-column tile
        h = N // 2
        il = lambda v: torch.cat([v[:h].view(h // 64, 64, *v.shape[1:]), v[h:].view(h // 64, 64, *v.shape[1:])], 1).reshape(v.shape).contiguous()
        wf, s_n, t_n = il(wf), il(s_n), il(t_n)
    No = N // 2 if act == 2 else N
    out = torch.full((M, No), float("nan"), device=DEV, dtype=torch.float16)
    L.check(lib.dfb_gemm_ln(L.ptr(x16), L.ptr(wf), M, N, C, L.ptr(t_n), L.ptr(s_n), L.ptr(st.contiguous()), tiles.value,
                            1e-5, act, None, L.ptr(out), sp_c, L.cur_stream()), "dfb_gemm_ln")
    sync()
    y = F.layer_norm(x_ref, (C,), gamma, beta, 1e-5) @ w.t() + b
    if act == 2:
        v, gate = y.chunk(2, dim=-1)
        y = v * F.gelu(gate)
    err = rel_l2(out.float(), y)
    print(f"\n[ln-fold] M={M} C={C} N={N} act={act}: rel-L2 {err:.2e}")
    assert torch.isfinite(out.float()).all() and err < 1.5e-3


def conv3x3(a16_nhwc, w_oihw16, bias=None, rowvec=None, residual=None, splits=0):
    B, H, W, C = a16_nhwc.shape
    N = w_oihw16.shape[0]
    wp = w_oihw16.permute(0, 2, 3, 1).reshape(N, 9 * C).contiguous()  # k = (ky*3+kx)*C + c
    out = torch.full((B, H, W, N), float("nan"), device=DEV, dtype=torch.float32)
    L.check(L.lib().dfb_conv3x3(L.ptr(a16_nhwc), L.ptr(wp), B, H, W, C, N, L.ptr(bias), L.ptr(rowvec),
                                L.ptr(residual), 0, L.ptr(out), None, splits, L.cur_stream()), "dfb_conv3x3")
    sync()
    return out


# the four UNet resolutions, B_eff = 2 (ragged last batch tile at 2x8) and an odd batch
@pytest.mark.parametrize("B,H,W,C,N,splits", [
    (2, 16, 64, 320, 320, 0), (2, 8, 32, 640, 640, 0), (2, 4, 16, 1280, 1280, 0),
    (2, 2, 8, 1280, 1280, 0), (3, 2, 8, 128, 64, 1), (1, 16, 64, 64, 128, 1), (5, 4, 16, 64, 64, 2),
])
def test_conv3x3(B, H, W, C, N, splits):
    g = torch.Generator(device="cpu").manual_seed(B + H + C + N)
    a = torch.randn(B, H, W, C, generator=g).to(DEV).half()
    w = (torch.randn(N, C, 3, 3, generator=g) / math.sqrt(9 * C)).to(DEV).half()
    bias = torch.randn(N, generator=g).to(DEV)
    rowvec = torch.randn(B, N, generator=g).to(DEV)
    res = torch.randn(B, H, W, N, generator=g).to(DEV)
    ref = F.conv2d(a.float().permute(0, 3, 1, 2), w.float(), bias, padding=1)
    ref = (ref + rowvec[:, :, None, None]).permute(0, 2, 3, 1) + res
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        ref = F.conv2d(a.float().permute(0, 3, 1, 2), w.float(), bias, padding=1)
        ref = (ref + rowvec[:, :, None, None]).permute(0, 2, 3, 1) + res
    finally:
        torch.backends.cudnn.allow_tf32 = old
    out = conv3x3(a, w, bias, rowvec, res, splits)
    assert torch.isfinite(out).all()
    assert rel_l2(out, ref) < 3e-6


# every (C, HW) GroupNorm call site of one UNet forward (SURVEY 8a'), with the source split the plan uses (the
# skip concat is read from two tensors), both eps (ResBlock 1e-5 + SiLU / SpatialTransformer.norm 1e-6, no
# SiLU), at B_eff = 2 and -- for the shapes whose launch geometry changes -- B_eff = 16
GN_SITES = [
    (320, 0, 1024, 1e-5, 1), (320, 0, 1024, 1e-6, 0), (320, 320, 1024, 1e-5, 1), (640, 320, 1024, 1e-5, 1),
    (320, 0, 256, 1e-5, 1), (640, 0, 256, 1e-5, 1), (640, 0, 256, 1e-6, 0), (640, 320, 256, 1e-5, 1),
    (640, 640, 256, 1e-5, 1), (1280, 640, 256, 1e-5, 1), (640, 0, 64, 1e-5, 1), (1280, 0, 64, 1e-5, 1),
    (1280, 0, 64, 1e-6, 0), (1280, 640, 64, 1e-5, 1), (1280, 1280, 64, 1e-5, 1), (1280, 0, 16, 1e-5, 1),
    (1280, 0, 16, 1e-6, 0), (1280, 1280, 16, 1e-5, 1), (128, 0, 1024, 1e-5, 1),
]


@pytest.mark.parametrize("B", [2, 16])
@pytest.mark.parametrize("C0,C1,HW,eps,silu", GN_SITES)
def test_groupnorm(C0, C1, HW, eps, silu, B):
    if B == 16 and (C0 + C1, HW) not in ((320, 1024), (960, 1024), (640, 256), (1920, 256), (1280, 64), (2560, 16)):
        pytest.skip("B_eff = 16 is exercised on one shape per level / source split")
    g = torch.Generator(device="cpu").manual_seed(C0 + C1 + HW)
    x0 = (torch.randn(B, HW, C0, generator=g) * 2 + 0.5).to(DEV)
    x1 = (torch.randn(B, HW, C1, generator=g) * 3 - 1).to(DEV) if C1 else None
    C = C0 + C1
    gamma = torch.randn(C, generator=g).to(DEV)
    beta = torch.randn(C, generator=g).to(DEV)
    out = torch.empty(B, HW, C, device=DEV, dtype=torch.float16)
    raw = torch.empty_like(out)
    L.check(L.lib().dfb_groupnorm(L.ptr(x0), C0, L.ptr(x1), C1, B, HW, L.ptr(gamma), L.ptr(beta), eps,
                                  silu, L.ptr(out), L.ptr(raw), L.cur_stream()), "dfb_groupnorm")
    sync()
    x = torch.cat([x0, x1], dim=-1) if C1 else x0
    ref = F.group_norm(x.permute(0, 2, 1), 32, gamma, beta, eps).permute(0, 2, 1)
    if silu:
        ref = F.silu(ref)
    assert rel_l2(out.float(), ref) < 6e-4  # fp16 output rounding (2^-11 per element)
    assert rel_l2(raw.float(), x) < 6e-4


@pytest.mark.parametrize("rows,C", [(2048, 320), (512, 640), (128, 1280), (32, 1280), (7, 256),
                                    (16384, 320), (4097, 640), (4096, 384), (8192, 1280)])
def test_layernorm(rows, C):
    g = torch.Generator(device="cpu").manual_seed(rows + C)
    x = (torch.randn(rows, C, generator=g) * 1.7 + 0.3).to(DEV)
    gamma = torch.randn(C, generator=g).to(DEV)
    beta = torch.randn(C, generator=g).to(DEV)
    out = torch.empty(rows, C, device=DEV, dtype=torch.float16)
    L.check(L.lib().dfb_layernorm(L.ptr(x), rows, C, L.ptr(gamma), L.ptr(beta), 1e-5, L.ptr(out),
                                  L.cur_stream()), "dfb_layernorm")
    sync()
    ref = F.layer_norm(x, (C,), gamma, beta, 1e-5)
    assert rel_l2(out.float(), ref) < 6e-4


@pytest.mark.parametrize("B,heads,Lq,Lk,d", [
    (2, 8, 1024, 1024, 40), (2, 8, 256, 256, 80), (2, 8, 64, 64, 160), (2, 8, 16, 16, 160),
    (2, 8, 1024, 32, 40), (2, 8, 64, 32, 160), (3, 8, 256, 33, 32), (1, 4, 100, 70, 16),
])
def test_attention(B, heads, Lq, Lk, d):
    """softmax(q k^T d^-0.5) v (attention_openai.py:170-193); q/k/v head-padded to dpad columns."""
    g = torch.Generator(device="cpu").manual_seed(Lq + Lk + d)
    dpad = (d + 15) // 16 * 16
    q = torch.zeros(B, Lq, heads, dpad)
    k = torch.zeros(B, Lk, heads, dpad)
    v = torch.zeros(B, Lk, heads, dpad)
    q[..., :d] = torch.randn(B, Lq, heads, d, generator=g)
    k[..., :d] = torch.randn(B, Lk, heads, d, generator=g)
    v[..., :d] = torch.randn(B, Lk, heads, d, generator=g)
    q, k, v = (t.to(DEV).half() for t in (q, k, v))
    out = torch.empty(B, Lq, heads * d, device=DEV, dtype=torch.float16)
    ld = heads * dpad
    L.check(L.lib().dfb_attention(L.ptr(q), ld, L.ptr(k), ld, L.ptr(v), ld, L.ptr(out), heads * d, B, heads,
                                  Lq, Lk, d, dpad, d ** -0.5, L.cur_stream()), "dfb_attention")
    sync()
    qf, kf, vf = (t.float()[..., :d].permute(0, 2, 1, 3) for t in (q, k, v))
    sim = torch.einsum("bhid,bhjd->bhij", qf, kf) * d ** -0.5
    ref = torch.einsum("bhij,bhjd->bhid", sim.softmax(-1), vf).permute(0, 2, 1, 3).reshape(B, Lq, heads * d)
    assert rel_l2(out.float(), ref) < 2e-3  # fp16 P and output rounding


def test_temb_and_small_ops():
    B, dim = 4, 320
    for t, is_float in ((torch.tensor([961, 1, 500, 41]), 0), (torch.tensor([961.0, 0.5, 333.25, 12.0]), 1)):
        td = t.to(DEV)
        out = torch.empty(B, dim, device=DEV, dtype=torch.float16)
        L.check(L.lib().dfb_temb(L.ptr(td), is_float, B, dim, L.ptr(out), L.cur_stream()), "dfb_temb")
        sync()
        half = dim // 2
        freqs = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32) / half).to(DEV)
        args = td[:, None].float() * freqs[None]
        ref = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
        assert (out.float() - ref).abs().max() < 1.5e-3  # fp16 rounding of values in [-1, 1] + fp32 sin/cos at ~1e3 rad
    g = torch.Generator(device="cpu").manual_seed(3)
    x = torch.randn(2, 4, 8, 64, generator=g).to(DEV)
    up = torch.empty(2, 8, 16, 64, device=DEV, dtype=torch.float16)
    L.check(L.lib().dfb_upsample2x_f16(L.ptr(x), L.ptr(up), 2, 4, 8, 64, L.cur_stream()), "upsample")
    ref = F.interpolate(x.permute(0, 3, 1, 2), scale_factor=2, mode="nearest").permute(0, 2, 3, 1).half()
    sync()
    assert torch.equal(up, ref)
    col = torch.empty(2 * 2 * 4, 9 * 64, device=DEV, dtype=torch.float16)
    L.check(L.lib().dfb_im2col_s2(L.ptr(x), L.ptr(col), 2, 4, 8, 64, L.cur_stream()), "im2col")
    sync()
    unf = F.unfold(x.permute(0, 3, 1, 2), 3, padding=1, stride=2)  # [B, C*9, L], k = c*9+tap
    unf = unf.view(2, 64, 9, 8).permute(0, 3, 2, 1).reshape(16, 9 * 64).half()
    assert torch.equal(col, unf)


def test_ddim_step_bit_exact():
    """CFG combine + DDIM update (ddim.py:241-245, 258-273) is bit-exact vs the same fp32 torch ops."""
    g = torch.Generator(device="cpu").manual_seed(9)
    n = 2 * 4 * 16 * 64
    x, eu, ec = (torch.randn(n, generator=g).to(DEV) for _ in range(3))
    a_t, a_prev, s = 0.0047, 0.0123, 4.5
    c1, c2, c3, c4 = math.sqrt(1 - a_t), math.sqrt(a_t), math.sqrt(a_prev), math.sqrt(1 - a_prev)
    xp, p0 = torch.empty_like(x), torch.empty_like(x)
    L.check(L.lib().dfb_ddim_step(L.ptr(x), L.ptr(eu), L.ptr(ec), None, s, c1, c2, c3, c4, 0.0, L.ptr(xp),
                                  L.ptr(p0), n, L.cur_stream()), "ddim_step")
    sync()
    f = lambda v: torch.tensor(v, dtype=torch.float32, device=DEV)
    e = eu + f(s) * (ec - eu)
    r0 = (x - f(c1) * e) / f(c2)
    rp = f(c3) * r0 + f(c4) * e
    assert torch.equal(p0, r0) and torch.equal(xp, rp)


# every (tile width, ring depth, K-splits) candidate the measured plan table (igemm_tuned.inc) may pick:
# ragged M (rows outside the tensor are neither exchanged nor stored), N that ends mid-tile, split
# factors that do not divide 128 rows, residual + bias epilogue; split-K results are bit-reproducible
@pytest.mark.parametrize("M,N,K", [(32, 1280, 1280), (200, 320, 640), (512, 640, 1920)])
@pytest.mark.parametrize("bn,deep", [(64, 0), (64, 1), (128, 1)])
def test_gemm_forced_plans(M, N, K, bn, deep):
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).to(DEV).half()
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(DEV).half()
    bias = torch.randn(N, generator=g).to(DEV)
    res = torch.randn(M, N, generator=g).to(DEV)
    ref = a.float() @ w.float().t() + bias + res
    lib = L.lib()
    try:
        lib.dfb_debug_igemm_force(bn, deep)
        for splits in (1, 2, 3, 5, 7, 8):
            out = gemm(a, w, bias=bias, residual=res, splits=splits)
            assert rel_l2(out, ref) < 2e-6, (bn, deep, splits)
            if splits > 1:
                again = gemm(a, w, bias=bias, residual=res, splits=splits)
                assert torch.equal(out, again), "split-K reduction must be deterministic"
    finally:
        lib.dfb_debug_igemm_force(0, -1)


@pytest.mark.parametrize("splits", [3, 6])
def test_conv3x3_split_ragged_batch(splits):
    """3x3 conv at the 2x8 level with B = 3 (the 128-row tile holds 8 samples: 5 of them do not exist)."""
    g = torch.Generator(device="cpu").manual_seed(17 + splits)
    B, H, W, C, N = 3, 2, 8, 256, 320
    a = torch.randn(B, H, W, C, generator=g).to(DEV).half()
    w = (torch.randn(N, C, 3, 3, generator=g) / math.sqrt(9 * C)).to(DEV).half()
    bias = torch.randn(N, generator=g).to(DEV)
    rowvec = torch.randn(B, N, generator=g).to(DEV)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        ref = F.conv2d(a.float().permute(0, 3, 1, 2), w.float(), bias, padding=1)
        ref = (ref + rowvec[:, :, None, None]).permute(0, 2, 3, 1)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    lib = L.lib()
    try:
        for bn, deep in ((64, 0), (128, 1)):
            lib.dfb_debug_igemm_force(bn, deep)
            out = conv3x3(a, w, bias, rowvec, None, splits)
            assert torch.isfinite(out).all()
            assert rel_l2(out, ref) < 3e-6, (bn, deep)
    finally:
        lib.dfb_debug_igemm_force(0, -1)


@pytest.mark.parametrize("B,H,W,C,C2,N,splits", [(2, 16, 64, 320, 640, 320, 0), (2, 4, 16, 128, 64, 192, 3),
                                                  (3, 2, 8, 64, 128, 64, 1)])
def test_conv3x3_with_fused_skip(B, H, W, C, C2, N, splits):
    """conv3x3(a) + conv1x1(a2) + both biases as one implicit GEMM (ResBlock conv2 + skip_connection)."""
    g = torch.Generator(device="cpu").manual_seed(B + C + C2 + N)
    a = torch.randn(B, H, W, C, generator=g).to(DEV).half()
    a2 = torch.randn(B, H, W, C2, generator=g).to(DEV).half()
    w = (torch.randn(N, C, 3, 3, generator=g) / math.sqrt(9 * C)).to(DEV).half()
    w2 = (torch.randn(N, C2, generator=g) / math.sqrt(C2)).to(DEV).half()
    b1 = torch.randn(N, generator=g).to(DEV)
    b2 = torch.randn(N, generator=g).to(DEV)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        ref = F.conv2d(a.float().permute(0, 3, 1, 2), w.float(), b1, padding=1).permute(0, 2, 3, 1)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    ref = ref + a2.float() @ w2.float().t() + b2
    wp = torch.cat([w.permute(0, 2, 3, 1).reshape(N, 9 * C), w2], dim=1).contiguous()
    out = torch.full((B, H, W, N), float("nan"), device=DEV, dtype=torch.float32)
    L.check(L.lib().dfb_conv3x3_cat(L.ptr(a), L.ptr(a2), C2, L.ptr(wp), B, H, W, C, N, L.ptr(b1), L.ptr(b2), None,
                                    L.ptr(out), None, splits, L.cur_stream()), "dfb_conv3x3_cat")
    sync()
    assert torch.isfinite(out).all()
    assert rel_l2(out, ref) < 3e-6


@pytest.mark.parametrize("C,HW,B", [(128, 65536, 1), (256, 16384, 2), (512, 16384, 1)])
def test_groupnorm_big_slab(C, HW, B):
    """Group slabs beyond one cluster's registers (first-stage decoder levels): statistics + apply path."""
    g = torch.Generator(device="cpu").manual_seed(C + HW)
    x = (torch.randn(B, HW, C, generator=g) * 1.7 + 0.4).to(DEV)
    gamma = (1 + 0.2 * torch.randn(C, generator=g)).to(DEV)
    beta = (0.1 * torch.randn(C, generator=g)).to(DEV)
    out = torch.empty(B, HW, C, device=DEV, dtype=torch.float16)
    raw = torch.empty_like(out)
    L.check(L.lib().dfb_groupnorm(L.ptr(x), C, None, 0, B, HW, L.ptr(gamma), L.ptr(beta), 1e-6, 1, L.ptr(out),
                                  L.ptr(raw), L.cur_stream()), "dfb_groupnorm")
    sync()
    ref = F.silu(F.group_norm(x.permute(0, 2, 1), 32, gamma, beta, 1e-6)).permute(0, 2, 1)
    assert rel_l2(out.float(), ref) < 1e-3  # fp16 output rounding
    assert rel_l2(raw.float(), x) < 1e-3
    again = torch.empty_like(out)
    L.check(L.lib().dfb_groupnorm(L.ptr(x), C, None, 0, B, HW, L.ptr(gamma), L.ptr(beta), 1e-6, 1, L.ptr(again),
                                  None, L.cur_stream()), "dfb_groupnorm")
    sync()
    assert torch.equal(out, again)


def test_softmax_rows():
    g = torch.Generator(device="cpu").manual_seed(9)
    x = (torch.randn(1024, 1024, generator=g) * 30).to(DEV)
    out = torch.empty(1024, 1024, device=DEV, dtype=torch.float16)
    L.check(L.lib().dfb_softmax_rows(L.ptr(x), 1024, 1024, 512 ** -0.5, L.ptr(out), L.cur_stream()), "softmax")
    sync()
    ref = torch.softmax(x * 512 ** -0.5, dim=1)
    assert rel_l2(out.float(), ref) < 1e-3


def test_groupnorm_large_batch():
    """B_eff = 16 at the 16x64 level: the grid already fills the machine, minimal cluster size."""
    g = torch.Generator(device="cpu").manual_seed(77)
    B, HW, C = 16, 1024, 320
    x = (torch.randn(B, HW, C, generator=g) * 1.3 - 0.2).to(DEV)
    gamma = (1 + 0.2 * torch.randn(C, generator=g)).to(DEV)
    beta = (0.1 * torch.randn(C, generator=g)).to(DEV)
    out = torch.empty(B, HW, C, device=DEV, dtype=torch.float16)
    L.check(L.lib().dfb_groupnorm(L.ptr(x), C, None, 0, B, HW, L.ptr(gamma), L.ptr(beta), 1e-5, 1, L.ptr(out),
                                  None, L.cur_stream()), "dfb_groupnorm")
    sync()
    ref = F.silu(F.group_norm(x.permute(0, 2, 1), 32, gamma, beta, 1e-5)).permute(0, 2, 1)
    assert rel_l2(out.float(), ref) < 1e-3


@pytest.mark.parametrize("M,N,K,act", [(512, 256, 128, 0), (2048, 320, 320, 0), (300, 128, 64, 0), (4096, 640, 640, 0),
                                       (1024, 2560, 320, 2), (1152, 384, 1600, 0)])
def test_gemm_cta_pair_mode(M, N, K, act):
    """tcgen05.mma.cta_group::2: two CTAs of a cluster share one 256 x 128 tile (each loads its 128 rows of A and half
    of the weight tile, both CTAs' TMA loads count on the leader's barrier, the commits are multicast).  Same result
    as the single-CTA kernel, bit for bit (identical K order and epilogue); odd M-tile counts leave the last pair's
    second CTA without rows."""
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).to(DEV).half()
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(DEV).half()
    bias = torch.randn(N, generator=g).to(DEV)
    res = torch.randn(M, N, generator=g).to(DEV) if act == 0 else None
    outs = []
    try:
        for pair in (0, 1):
            L.lib().dfb_debug_igemm_pair(pair)
            o32 = torch.full((M, N), float("nan"), device=DEV) if act == 0 else None
            o16 = torch.full((M, N // 2), float("nan"), device=DEV, dtype=torch.float16) if act == 2 else None
            L.check(L.lib().dfb_gemm(L.ptr(a), L.ptr(w), M, N, K, L.ptr(bias), L.ptr(res), act, L.ptr(o32), L.ptr(o16), 1,
                                     L.cur_stream()), "dfb_gemm")
            sync()
            outs.append(o32 if act == 0 else o16)
    finally:
        L.lib().dfb_debug_igemm_pair(-1)
    assert torch.isfinite(outs[1].float()).all()
    if act == 0:
        ref = a.float() @ w.float().t() + bias + res
        assert rel_l2(outs[1], ref) < 3e-6
    assert torch.equal(outs[0], outs[1])


# the weight-streaming layers (M <= 32: three quarters or more of the tile rows lie outside the tensor and may not leak
# into a sum, an exchange or a store): M = 2 (embedding Linears), a ragged M, the level-3 conv with the fused
# skip, and the GEGLU epilogue, with the 128-wide tile and deep ring the planner gives them
@pytest.mark.parametrize("M", [2, 17, 32])
@pytest.mark.parametrize("splits", [1, 3, 8])
def test_gemm_small_m_box(M, splits):
    N, K = 1280, 3840
    g = torch.Generator(device="cpu").manual_seed(M * 31 + splits)
    a = torch.randn(M, K, generator=g).to(DEV).half()
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(DEV).half()
    bias = torch.randn(N, generator=g).to(DEV)
    res = torch.randn(M, N, generator=g).to(DEV)
    ref = a.float() @ w.float().t() + bias + res
    lib = L.lib()
    try:
        lib.dfb_debug_igemm_force(128, 1)
        out = gemm(a, w, bias=bias, residual=res, splits=splits)
        again = gemm(a, w, bias=bias, residual=res, splits=splits)
    finally:
        lib.dfb_debug_igemm_force(0, -1)
    assert torch.isfinite(out).all()
    assert rel_l2(out, ref) < 4e-6, rel_l2(out, ref)   # (few elements at M = 2: a noisier statistic than the big shapes)
    assert torch.equal(out, again)


def test_gemm_small_m_box_geglu():
    g = torch.Generator(device="cpu").manual_seed(11)
    M, C = 32, 1280
    a = torch.randn(M, C, generator=g).to(DEV).half()
    w = (torch.randn(8 * C, C, generator=g) / math.sqrt(C)).to(DEV).half()
    bias = torch.randn(8 * C, generator=g).to(DEV)
    x, gate = (a.float() @ w.float().t() + bias).chunk(2, dim=-1)
    ref = x * F.gelu(gate)
    C4 = 4 * C
    wi = torch.cat([w[:C4].view(C4 // 64, 64, C), w[C4:].view(C4 // 64, 64, C)], dim=1).reshape(8 * C, C).contiguous()
    bi = torch.cat([bias[:C4].view(-1, 64), bias[C4:].view(-1, 64)], dim=1).reshape(-1).contiguous()
    out = gemm(a, wi, bias=bi, act=2, out_dtype=torch.float16, splits=1)
    assert rel_l2(out.float(), ref) < 1e-3


@pytest.mark.parametrize("B,splits", [(2, 0), (2, 5), (1, 8)])
def test_conv3x3_small_m_box_with_skip(B, splits):
    H, W, C, C2, N = 2, 8, 1280, 1280, 1280
    g = torch.Generator(device="cpu").manual_seed(B + splits)
    a = torch.randn(B, H, W, C, generator=g).to(DEV).half()
    a2 = torch.randn(B, H, W, C2, generator=g).to(DEV).half()
    w = (torch.randn(N, C, 3, 3, generator=g) / math.sqrt(9 * C)).to(DEV).half()
    w2 = (torch.randn(N, C2, generator=g) / math.sqrt(C2)).to(DEV).half()
    b1 = torch.randn(N, generator=g).to(DEV)
    b2 = torch.randn(N, generator=g).to(DEV)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        ref = F.conv2d(a.float().permute(0, 3, 1, 2), w.float(), b1, padding=1).permute(0, 2, 3, 1)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    ref = ref + a2.float() @ w2.float().t() + b2
    wp = torch.cat([w.permute(0, 2, 3, 1).reshape(N, 9 * C), w2], dim=1).contiguous()
    out = torch.full((B, H, W, N), float("nan"), device=DEV, dtype=torch.float32)
    L.check(L.lib().dfb_conv3x3_cat(L.ptr(a), L.ptr(a2), C2, L.ptr(wp), B, H, W, C, N, L.ptr(b1), L.ptr(b2), None,
                                    L.ptr(out), None, splits, L.cur_stream()), "dfb_conv3x3_cat")
    sync()
    assert torch.isfinite(out).all()
    assert rel_l2(out, ref) < 3e-6


# wide tiles (128 x bn_run, bn_run = any multiple of 16 up to 256, chosen so that the tiles cover N without padding):
# the large-M variant.  Plain / residual / SiLU / fp16 epilogues, N that is not a multiple of the tile, ragged M,
# the 3x3 conv with per-sample vector, and the LayerNorm-fold producer + consumer pair (statistics of the two
# column halves of a tile are merged by the thread that owns the row)
@pytest.mark.parametrize("M,N,K", [(2048, 320, 320), (4096, 640, 1280), (1024, 1280, 640), (300, 384, 128), (512, 1152, 320),
                                   (256, 144, 64), (640, 2560, 192), (16384, 320, 1600)])
def test_gemm_wide_tiles(M, N, K):
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).to(DEV).half()
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(DEV).half()
    bias = torch.randn(N, generator=g).to(DEV)
    res = torch.randn(M, N, generator=g).to(DEV)
    lib = L.lib()
    try:
        lib.dfb_debug_igemm_force(256, 1)
        out = gemm(a, w, bias=bias, residual=res)
        out16 = gemm(a, w, bias=bias, act=1, out_dtype=torch.float16)
        base = gemm(a, w, bias=bias, residual=res, splits=1) if False else None
    finally:
        lib.dfb_debug_igemm_force(0, -1)
    ref = a.float() @ w.float().t() + bias
    assert torch.isfinite(out).all() and rel_l2(out, ref + res) < 2e-6
    assert rel_l2(out16.float(), F.silu(ref)) < 6e-4


@pytest.mark.parametrize("B,H,W,C,N", [(2, 16, 64, 320, 320), (4, 8, 32, 128, 640), (3, 4, 16, 64, 192)])
def test_conv3x3_wide_tiles(B, H, W, C, N):
    g = torch.Generator(device="cpu").manual_seed(B + H + C + N)
    a = torch.randn(B, H, W, C, generator=g).to(DEV).half()
    w = (torch.randn(N, C, 3, 3, generator=g) / math.sqrt(9 * C)).to(DEV).half()
    bias = torch.randn(N, generator=g).to(DEV)
    rowvec = torch.randn(B, N, generator=g).to(DEV)
    res = torch.randn(B, H, W, N, generator=g).to(DEV)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        ref = F.conv2d(a.float().permute(0, 3, 1, 2), w.float(), bias, padding=1)
        ref = (ref + rowvec[:, :, None, None]).permute(0, 2, 3, 1) + res
    finally:
        torch.backends.cudnn.allow_tf32 = old
    lib = L.lib()
    try:
        lib.dfb_debug_igemm_force(256, 1)
        out = conv3x3(a, w, bias, rowvec, res, 0)
    finally:
        lib.dfb_debug_igemm_force(0, -1)
    assert torch.isfinite(out).all() and rel_l2(out, ref) < 3e-6


@pytest.mark.parametrize("M,C,N", [(2048, 320, 1152), (512, 640, 1920), (300, 192, 384)])
def test_gemm_layernorm_fold_wide_tiles(M, C, N):
    lib = L.lib()
    try:
        lib.dfb_debug_igemm_force(256, 1)
        test_gemm_layernorm_fold(M, C, N, 0, 0, 0)
    finally:
        lib.dfb_debug_igemm_force(0, -1)
