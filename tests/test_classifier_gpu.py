"""Row a15 / N2 on the GPU: the double-guidance classifier's probability and d/dx log p through the hand-written
forward + backward (diff_foley_b200/classifier.py::_Native over the C ABI) against the reference's own
Classifier_Backbone + DDIMSampler.cal_classifier_loglikelihood_grad (tests/golden/classifier_*.npz), and the
backward kernels one by one against torch autograd."""
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from diff_foley_b200 import _lib as L
from diff_foley_b200.classifier import AlignmentClassifierDoubleGuidanceB200
from oracle import classifier_oracle

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DEV = "cuda"
CLF_SMALL = dict(classifier_oracle.DIFF_FOLEY_CLASSIFIER, model_channels=64, num_heads=4, context_dim=64)


def rel_l2(a, b):
    a, b = torch.as_tensor(a).double().flatten().cpu(), torch.as_tensor(b).double().flatten().cpu()
    return float((a - b).norm() / b.norm())


def build(cfg, seed):
    params = dict(image_size=32, in_channels=4, out_channels=1, model_channels=cfg["model_channels"],
                  attention_resolutions=list(cfg["attention_resolutions"]), num_res_blocks=1,
                  channel_mult=list(cfg["channel_mult"]), num_heads=cfg["num_heads"], use_spatial_transformer=True,
                  transformer_depth=1, context_dim=cfg["context_dim"], use_checkpoint=True, legacy=False)
    clf = AlignmentClassifierDoubleGuidanceB200(params, cond_stage_params=dict(origin_dim=64, embed_dim=64, seq_len=40))
    clf.model.load_state_dict(classifier_oracle.seeded_state_dict(cfg, seed))
    return clf.cuda()


@pytest.mark.parametrize("name,cfg", [("classifier_small", CLF_SMALL), ("classifier_full", classifier_oracle.DIFF_FOLEY_CLASSIFIER)])
def test_classifier_grad_matches_reference(name, cfg):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    clf = build(cfg, int(g["seed"]))
    x, t, f = (torch.from_numpy(g[k]).cuda() for k in ("x", "t", "feats"))
    prob = clf.probability(x, t, f)
    grad = clf.loglikelihood_grad(x, t, f, 50.0)
    torch.cuda.synchronize()
    ep, eg = rel_l2(prob, g["prob"]), rel_l2(grad, g["grad"])
    print(f"\n[parity] {name}: prob rel-L2 {ep:.3e}, grad_x log p (scale 50) rel-L2 {eg:.3e}")
    assert torch.isfinite(grad).all() and grad.shape == x.shape
    # fp16 operands through ~45 forward + ~45 backward GEMMs (the UNet's single forward measures 9e-4)
    assert ep < 1e-3 and eg < 2.5e-3
    again = clf.loglikelihood_grad(x, t, f, 50.0)
    assert torch.equal(grad, again), "the gradient must be bit-reproducible (no atomics anywhere)"


def test_groupnorm_bwd_matches_autograd():
    g = torch.Generator().manual_seed(1)
    for (C, HW, silu, eps) in [(128, 1024, 1, 1e-5), (256, 256, 0, 1e-6), (256, 64, 1, 1e-5), (64, 16, 1, 1e-5)]:
        B = 3
        x = (torch.randn(B, HW, C, generator=g) * 1.5 + 0.3).to(DEV).requires_grad_(True)
        gam, bet = torch.randn(C, generator=g).to(DEV), torch.randn(C, generator=g).to(DEV)
        dy = torch.randn(B, HW, C, generator=g).to(DEV)
        add = torch.randn(B, HW, C, generator=g).to(DEV)
        y = F.group_norm(x.permute(0, 2, 1), 32, gam, bet, eps).permute(0, 2, 1)
        if silu:
            y = F.silu(y)
        (ref,) = torch.autograd.grad(y, x, dy)
        d32 = torch.empty(B, HW, C, device=DEV)
        d16 = torch.empty(B, HW, C, device=DEV, dtype=torch.float16)
        L.check(L.lib().dfb_groupnorm_bwd(L.ptr(x.detach()), C, B, HW, L.ptr(gam), L.ptr(bet), eps, silu, L.ptr(dy), L.ptr(add),
                                          L.ptr(d32), L.ptr(d16), L.cur_stream()), "dfb_groupnorm_bwd")
        torch.cuda.synchronize()
        assert rel_l2(d32, ref + add) < 2e-5 and rel_l2(d16.float(), ref + add) < 6e-4


def test_layernorm_bwd_matches_autograd():
    g = torch.Generator().manual_seed(2)
    for rows, C in [(2048, 256), (64, 256), (7, 64)]:
        x = (torch.randn(rows, C, generator=g) * 1.3 - 0.2).to(DEV).requires_grad_(True)
        gam, bet = torch.randn(C, generator=g).to(DEV), torch.randn(C, generator=g).to(DEV)
        dy, add = torch.randn(rows, C, generator=g).to(DEV), torch.randn(rows, C, generator=g).to(DEV)
        (ref,) = torch.autograd.grad(F.layer_norm(x, (C,), gam, bet, 1e-5), x, dy)
        d32, d16 = torch.empty(rows, C, device=DEV), torch.empty(rows, C, device=DEV, dtype=torch.float16)
        L.check(L.lib().dfb_layernorm_bwd(L.ptr(x.detach()), rows, C, L.ptr(gam), 1e-5, L.ptr(dy), L.ptr(add), L.ptr(d32),
                                          L.ptr(d16), L.cur_stream()), "dfb_layernorm_bwd")
        torch.cuda.synchronize()
        assert rel_l2(d32, ref + add) < 2e-5 and rel_l2(d16.float(), ref + add) < 6e-4


@pytest.mark.parametrize("B,heads,Lq,Lk,d", [(2, 8, 256, 256, 32), (3, 4, 64, 33, 16), (1, 8, 64, 64, 32), (2, 4, 100, 70, 64)])
def test_attention_bwd_matches_autograd(B, heads, Lq, Lk, d):
    g = torch.Generator().manual_seed(B + Lq + Lk + d)
    C = heads * d
    q = (torch.randn(B * Lq, C, generator=g)).to(DEV).half()
    k = (torch.randn(B * Lk, C, generator=g)).to(DEV).half()
    v = (torch.randn(B * Lk, C, generator=g)).to(DEV).half()
    dO = torch.randn(B * Lq, C, generator=g).to(DEV)
    qf, kf, vf = (t.float().requires_grad_(True) for t in (q, k, v))
    sp = lambda t, n: t.reshape(B, n, heads, d).permute(0, 2, 1, 3)
    att = (torch.einsum("bhid,bhjd->bhij", sp(qf, Lq), sp(kf, Lk)) * d ** -0.5).softmax(-1)
    o = torch.einsum("bhij,bhjd->bhid", att, sp(vf, Lk)).permute(0, 2, 1, 3).reshape(B * Lq, C)
    rq, rk, rv = torch.autograd.grad(o, (qf, kf, vf), dO)
    o16 = o.detach().half()
    dq, dk, dv = (torch.zeros(n, C, device=DEV, dtype=torch.float16) for n in (B * Lq, B * Lk, B * Lk))
    ws = torch.empty(2, B * heads * Lq, device=DEV)
    L.check(L.lib().dfb_attention_bwd(L.ptr(q), C, L.ptr(k), C, L.ptr(v), C, L.ptr(o16), C, L.ptr(dO), C, B, heads, Lq, Lk, d,
                                      d ** -0.5, L.ptr(dq), C, L.ptr(dk), C, L.ptr(dv), C, L.ptr(ws[0]), L.ptr(ws[1]),
                                      L.cur_stream()), "dfb_attention_bwd")
    torch.cuda.synchronize()
    assert rel_l2(dq.float(), rq) < 2e-3 and rel_l2(dk.float(), rk) < 2e-3 and rel_l2(dv.float(), rv) < 2e-3


def test_geglu_and_col2im_match_autograd():
    g = torch.Generator().manual_seed(4)
    M, Fh = 300, 256
    proj = torch.randn(M, 2 * Fh, generator=g).to(DEV).requires_grad_(True)
    dh = torch.randn(M, Fh, generator=g).to(DEV)
    a, gate = proj.chunk(2, dim=-1)
    y = a * F.gelu(gate)
    (ref,) = torch.autograd.grad(y, proj, dh)
    h = torch.empty(M, Fh, device=DEV, dtype=torch.float16)
    dp = torch.empty(M, 2 * Fh, device=DEV, dtype=torch.float16)
    L.check(L.lib().dfb_geglu_fwd(L.ptr(proj.detach()), M, Fh, L.ptr(h), L.cur_stream()), "geglu_fwd")
    L.check(L.lib().dfb_geglu_bwd(L.ptr(proj.detach()), L.ptr(dh), M, Fh, L.ptr(dp), L.cur_stream()), "geglu_bwd")
    torch.cuda.synchronize()
    assert rel_l2(h.float(), y) < 6e-4 and rel_l2(dp.float(), ref) < 6e-4
    # stride-2 conv backward-data: dcol = dY . W (fp32 here), gathered by col2im_s2 == conv_transpose of autograd
    B, H, W, C, N = 2, 8, 16, 64, 32
    x = torch.randn(B, C, H, W, generator=g).to(DEV).requires_grad_(True)
    w = (torch.randn(N, C, 3, 3, generator=g) / math.sqrt(9 * C)).to(DEV)
    dy = torch.randn(B, N, H // 2, W // 2, generator=g).to(DEV)
    (rx,) = torch.autograd.grad(F.conv2d(x, w, stride=2, padding=1), x, dy)
    wf = w.permute(0, 2, 3, 1).reshape(N, 9 * C)                       # [N, (ky*3+kx)*C + c]
    dcol = (dy.permute(0, 2, 3, 1).reshape(-1, N) @ wf).contiguous()   # [B*Ho*Wo, 9C]
    dx = torch.empty(B, H, W, C, device=DEV)
    L.check(L.lib().dfb_col2im_s2(L.ptr(dcol), B, H, W, C, None, L.ptr(dx), None, L.cur_stream()), "col2im_s2")
    torch.cuda.synchronize()
    assert rel_l2(dx.permute(0, 3, 1, 2), rx) < 1e-5
