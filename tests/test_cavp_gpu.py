"""CAVP encoders on the CUDA kernels vs the reference's CAVP_Inference outputs
(tests/golden/cavp_*.npz, produced by the reference through the mmcv shim).  fp16 activations through
53 (video) / 12 (audio) convolutions without any re-normalisation -> tolerance 5e-3 on the raw
features; the L2-normalised features the pipeline consumes are checked too."""
import os

import numpy as np
import pytest
import torch

from diff_foley_b200 import _lib as L
from diff_foley_b200.cavp import CAVPInferenceB200
from oracle import cavp_oracle

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel_l2(a, b):
    a, b = torch.as_tensor(a).double().flatten().cpu(), torch.as_tensor(b).double().flatten().cpu()
    return float((a - b).norm() / b.norm())


def inputs(g):
    B, T, HW, spec_T = (int(v) for v in g["shape"])
    gen = torch.Generator().manual_seed(int(g["seed"]) + 77)
    return torch.rand(B, T, 3, HW, HW, generator=gen), torch.randn(B, 128, spec_T, generator=gen)


def test_pool_and_im2col_kernels():
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(0)
    x = torch.randn(3, 10, 12, 16, generator=g).cuda().half()           # [NI,H,W,C]
    for k, s, p, is_max in (((3, 3), (2, 2), (1, 1), True), ((2, 2), (2, 2), (0, 0), False),
                            ((1, 2), (1, 2), (0, 0), False), ((3, 1), (1, 1), (1, 0), False),
                            ((3, 1), (1, 1), (1, 0), True)):
        Ho, Wo = (10 + 2 * p[0] - k[0]) // s[0] + 1, (12 + 2 * p[1] - k[1]) // s[1] + 1
        out = torch.empty(3, Ho, Wo, 16, device="cuda", dtype=torch.float16)
        L.check(L.lib().dfb_pool2d_f16(L.ptr(x), L.ptr(out), 3, 10, 12, 16, k[0], k[1], s[0], s[1], p[0], p[1],
                                       int(is_max), L.cur_stream()))
        xc = x.float().permute(0, 3, 1, 2)
        ref = F.max_pool2d(xc, k, s, p) if is_max else F.avg_pool2d(xc, k, s, p)
        torch.cuda.synchronize()
        assert rel_l2(out.float(), ref.permute(0, 2, 3, 1)) < 1e-3
    col = torch.empty(3 * 5 * 6, 9 * 16 + 48, device="cuda", dtype=torch.float16)
    L.check(L.lib().dfb_im2col_f16(L.ptr(x), L.ptr(col), 3, 10, 12, 16, 3, 3, 2, 1, 9 * 16 + 48, L.cur_stream()))
    torch.cuda.synchronize()
    unf = F.unfold(x.float().permute(0, 3, 1, 2), 3, padding=1, stride=2).view(3, 16, 9, 30).permute(0, 3, 2, 1)
    assert torch.equal(col[:, :144].float(), unf.reshape(90, 144)) and float(col[:, 144:].abs().max()) == 0


def test_conv_taps_temporal_and_residual():
    """(3,1,1) temporal conv + fp16 identity + ReLU (Bottleneck3d conv1 / conv3 epilogue)."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(1)
    B, T, H, W, C, N = 2, 8, 7, 7, 64, 128
    a = torch.randn(B, T, H, W, C, generator=g).cuda().half()
    w = (torch.randn(N, C, 3, 1, 1, generator=g) / (3 * C) ** 0.5).cuda().half()
    bias = torch.randn(N, generator=g).cuda()
    res = torch.randn(B, T, H, W, N, generator=g).cuda().half()
    wp = w.permute(0, 2, 3, 4, 1).reshape(N, 3 * C).contiguous()
    out = torch.empty(B, T, H, W, N, device="cuda", dtype=torch.float16)
    L.check(L.lib().dfb_conv_taps(L.ptr(a), L.ptr(wp), B, T, H, W, C, N, 3, 1, 1, L.ptr(bias), L.ptr(res), 3, None,
                                  L.ptr(out), 0, L.cur_stream()))
    torch.cuda.synchronize()
    ref = F.conv3d(a.float().permute(0, 4, 1, 2, 3), w.float(), bias, 1, (1, 0, 0)).permute(0, 2, 3, 4, 1)
    ref = F.relu(ref + res.float())
    assert rel_l2(out.float(), ref) < 1e-3


@pytest.mark.parametrize("name", ["cavp_small", "cavp_full"])
def test_cavp_matches_reference(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    m = CAVPInferenceB200()
    m.load_state_dict(cavp_oracle.seeded_state_dict(int(g["seed"])), strict=False)
    m = m.cuda().eval()
    video, spec = inputs(g)
    v_raw = m.encode_video(video.cuda(), normalize=False, pool=False)
    v = m.encode_video(video.cuda(), normalize=True, pool=False)
    s_raw = m.encode_spec(spec.cuda(), normalize=False, pool=False)
    s = m.encode_spec(spec.cuda(), normalize=True, pool=False)
    torch.cuda.synchronize()
    errs = dict(video_raw=rel_l2(v_raw, g["video_raw"]), video=rel_l2(v, g["video_feat"]),
                spec_raw=rel_l2(s_raw, g["spec_raw"]), spec=rel_l2(s, g["spec_feat"]))
    print(f"\n[parity] {name}: " + ", ".join(f"{k} {e:.3e}" for k, e in errs.items()) + f"  ({m.launches} launches)")
    assert all(torch.isfinite(t).all() for t in (v_raw, s_raw))
    assert max(errs.values()) < 5e-3
    assert m.launches > 100
