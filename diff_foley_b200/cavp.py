"""CAVP video / audio encoders on the B200 kernels (BASELINE config 5).

`CAVPInferenceB200` mirrors `CAVP_Inference` (inference/model/cavp_model.py:9-96): same constructor,
same state-dict keys (`video_encoder.conv1.conv.weight`, `...bn.running_mean`, `spec_encoder.*`,
`video_project_head.*`, `logit_scale`; demo_util.py:107-121 loads them after stripping `module.`),
same `encode_video(video, normalize, pool)` / `encode_spec(spec, normalize, pool)` contracts.

Compute path (all convolutions / linears on `igemm_tcgen05_kernel`, fp16 channels-last
activations, fp32 accumulation):
  * BatchNorm is folded into the preceding conv once (w' = w*g/sqrt(var+eps), b' = beta - mu*g/sqrt(..)).
  * SlowOnly-R50 (cavp_modules.py:1233-1268 / ResNet3d :331-871 / Bottleneck3d :167-328):
    (1,1,1) convs are GEMMs over B*T*H*W rows, (1,3,3) convs are 9-tap and the (3,1,1) temporal convs
    3-tap implicit GEMMs (`dfb_conv_taps`, TMA zero fill = padding), the stride-2 convs and the
    (1,7,7)/2 stem go through a fp16 im2col (`dfb_im2col_f16`) into the same GEMM; the identity add +
    ReLU of every bottleneck is the conv3 epilogue (fp16 residual); max / average pooling are
    `dfb_pool2d_f16`.
  * PANNs Cnn14 (cavp_modules.py:1487-1546): twelve 3x3 conv+BN+ReLU as 9-tap implicit GEMMs, average
    pools, the time max+avg pooling, `fc1` applied twice (a reference quirk, :1543-1544) and
    `final_project` as GEMMs.
Only layout conversion at the boundary (NCTHW fp32 -> channels-last fp16, zero-padding the 3 / 1 input
channels to 8) and the final L2-normalise / MaxPool1d over a [B,T,512] tensor use torch ops.
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib as L

RELU, NONE = 3, 0


class _ConvBN3d(nn.Module):
    """Parameter holder named like mmcv's ConvModule: `.conv` (no bias) + `.bn`."""

    def __init__(self, cin, cout, k, stride=(1, 1, 1)):
        super().__init__()
        self.conv = nn.Conv3d(cin, cout, k, stride=stride, padding=tuple(x // 2 for x in k), bias=False)
        self.bn = nn.BatchNorm3d(cout)
        self.k, self.stride = tuple(k), tuple(stride)


class _Bottleneck(nn.Module):
    def __init__(self, inplanes, planes, spatial_stride, inflate, downsample):
        super().__init__()
        self.conv1 = _ConvBN3d(inplanes, planes, (3, 1, 1) if inflate else (1, 1, 1))
        self.conv2 = _ConvBN3d(planes, planes, (1, 3, 3), (1, spatial_stride, spatial_stride))
        self.conv3 = _ConvBN3d(planes, planes * 4, (1, 1, 1))
        if downsample:
            self.downsample = _ConvBN3d(inplanes, planes * 4, (1, 1, 1), (1, spatial_stride, spatial_stride))


class _SlowOnlyR50(nn.Module):
    """depth 50, stage blocks (3,4,6,3), spatial strides (1,2,2,2), no temporal stride,
    inflate (0,0,1,1) -- ResNet3dSlowOnly defaults (cavp_modules.py:1250-1266)."""

    def __init__(self):
        super().__init__()
        self.conv1 = _ConvBN3d(3, 64, (1, 7, 7), (1, 2, 2))
        inplanes = 64
        for i, (blocks, stride, inflate) in enumerate(zip((3, 4, 6, 3), (1, 2, 2, 2), (0, 0, 1, 1))):
            planes = 64 * 2 ** i
            layer = []
            for b in range(blocks):
                s = stride if b == 0 else 1
                layer.append(_Bottleneck(inplanes, planes, s, bool(inflate), b == 0 and (s != 1 or inplanes != planes * 4)))
                inplanes = planes * 4
            self.add_module(f"layer{i + 1}", nn.Sequential(*layer))


class _ConvBlock(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1, bias=False)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(cout)
        self.bn2 = nn.BatchNorm2d(cout)


class _Cnn14(nn.Module):
    def __init__(self, embed_dim):
        super().__init__()
        self.bn = nn.BatchNorm2d(128)
        chans = [1, 64, 128, 256, 512, 1024, 2048]
        for i in range(6):
            self.add_module(f"conv_block{i + 1}", _ConvBlock(chans[i], chans[i + 1]))
        self.fc1 = nn.Linear(2048, 2048)
        self.final_project = nn.Linear(2048, embed_dim)


def _fold(conv_w, bn):
    """conv weight [N,C,...] + BatchNorm (eval) -> folded weight, bias (fp32)."""
    scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    w = conv_w.detach().float() * scale.view(-1, *([1] * (conv_w.dim() - 1)))
    b = bn.bias.detach().float() - bn.running_mean.detach().float() * scale
    return w, b


def _pack(w, cpad=None, kpad=None):
    """[N,C,*k] -> fp16 [N, taps*Cp] with k = tap*Cp + c (tap row-major over the kernel dims)."""
    n, c = w.shape[:2]
    taps = int(np.prod(w.shape[2:])) if w.dim() > 2 else 1
    w = w.reshape(n, c, taps).permute(0, 2, 1)                       # [N, taps, C]
    if cpad and cpad > c:
        w = F.pad(w, (0, cpad - c))
    w = w.reshape(n, -1)
    if kpad and kpad > w.shape[1]:
        w = F.pad(w, (0, kpad - w.shape[1]))
    return w.half().contiguous()


class CAVPInferenceB200(nn.Module):
    def __init__(self, video_encode="Slowonly_pool", spec_encode="cnn14_pool", embed_dim=512,
                 video_pretrained=False, audio_pretrained=False):
        super().__init__()
        assert video_encode == "Slowonly_pool" and spec_encode == "cnn14_pool"
        self.video_encode, self.spec_encode = video_encode, spec_encode
        self.video_encoder = _SlowOnlyR50()
        self.video_project_head = nn.Linear(2048, embed_dim)
        self.video_pool = nn.MaxPool1d(kernel_size=16)
        self.spec_encoder = _Cnn14(embed_dim=512)
        self.spec_project_head = nn.Identity()
        self.spec_pool = nn.MaxPool1d(kernel_size=16)
        self.logit_scale = nn.Parameter(torch.ones([]) * math.log(1 / 0.07))
        self._packed = None
        self.launches = 0

    def _apply(self, fn, *a, **k):
        self._packed = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._packed = None
        return super().load_state_dict(*a, **k)

    # ------------------------------------------------------------------------------- packing
    def _pack_all(self, dev):
        if self._packed is not None:
            return self._packed
        P = {}
        ve = self.video_encoder
        w, b = _fold(ve.conv1.conv.weight[:, :, 0], ve.conv1.bn)           # [64,3,7,7]
        P["stem"] = (_pack(w, cpad=8, kpad=448).to(dev), b.to(dev))
        for li in range(1, 5):
            for bi, blk in enumerate(getattr(ve, f"layer{li}")):
                for name in ("conv1", "conv2", "conv3", "downsample"):
                    if not hasattr(blk, name):
                        continue
                    m = getattr(blk, name)
                    w, b = _fold(m.conv.weight, m.bn)
                    if name in ("conv2", "downsample") and m.stride[1] != 1:
                        w = w[:, :, 0]                                     # 2-D kernel for the im2col path
                    P[f"l{li}.{bi}.{name}"] = (_pack(w).to(dev), b.to(dev))
        P["vproj"] = (self.video_project_head.weight.detach().half().contiguous().to(dev),
                      self.video_project_head.bias.detach().float().to(dev))
        se = self.spec_encoder
        s0 = se.bn.weight.detach().float() / torch.sqrt(se.bn.running_var.detach().float() + se.bn.eps)
        P["spec_bn"] = (s0.to(dev), (se.bn.bias.detach().float() - se.bn.running_mean.detach().float() * s0).to(dev))
        for i in range(1, 7):
            cb = getattr(se, f"conv_block{i}")
            w1, b1 = _fold(cb.conv1.weight, cb.bn1)
            w2, b2 = _fold(cb.conv2.weight, cb.bn2)
            P[f"cb{i}.1"] = (_pack(w1, cpad=8, kpad=128).to(dev) if i == 1 else _pack(w1).to(dev), b1.to(dev))
            P[f"cb{i}.2"] = (_pack(w2).to(dev), b2.to(dev))
        P["fc1"] = (se.fc1.weight.detach().half().contiguous().to(dev), se.fc1.bias.detach().float().to(dev))
        P["final"] = (se.final_project.weight.detach().half().contiguous().to(dev),
                      se.final_project.bias.detach().float().to(dev))
        self._packed = P
        return P

    # --------------------------------------------------------------------------- kernel calls
    def _gemm(self, a, wb, act=NONE, out_dtype=torch.float16, residual=None):
        w, b = wb
        M, K = a.shape
        N = w.shape[0]
        out = torch.empty(M, N, device=a.device, dtype=out_dtype)
        if residual is None:
            L.check(L.lib().dfb_gemm(L.ptr(a), L.ptr(w), M, N, K, L.ptr(b), None, act,
                                     L.ptr(out) if out_dtype == torch.float32 else None,
                                     L.ptr(out) if out_dtype == torch.float16 else None, 0, L.cur_stream()), "dfb_gemm")
        else:  # fp16 identity path: the (1,1,1) conv3 of a bottleneck, as a 1-tap conv over M rows
            L.check(L.lib().dfb_conv_taps(L.ptr(a), L.ptr(w), 1, 1, 1, M, K, N, 1, 1, 1, L.ptr(b), L.ptr(residual),
                                          act, None, L.ptr(out), 0, L.cur_stream()), "dfb_conv_taps")
        self.launches += 1
        return out

    def _conv(self, a, shape5, wb, k, act):
        """a: fp16 [B,T,H,W,C] contiguous; stride-1 'same' conv with kernel k=(kt,kh,kw)."""
        B, T, H, W, C = shape5
        w, b = wb
        N = w.shape[0]
        out = torch.empty(B * T * H * W, N, device=a.device, dtype=torch.float16)
        L.check(L.lib().dfb_conv_taps(L.ptr(a), L.ptr(w), B, T, H, W, C, N, k[0], k[1], k[2], L.ptr(b), None, act,
                                      None, L.ptr(out), 0, L.cur_stream()), "dfb_conv_taps")
        self.launches += 1
        return out

    def _im2col(self, a, NI, H, W, C, kh, kw, stride, pad, kpad):
        Ho, Wo = (H + 2 * pad - kh) // stride + 1, (W + 2 * pad - kw) // stride + 1
        out = torch.empty(NI * Ho * Wo, kpad, device=a.device, dtype=torch.float16)
        L.check(L.lib().dfb_im2col_f16(L.ptr(a), L.ptr(out), NI, H, W, C, kh, kw, stride, pad, kpad, L.cur_stream()),
                "dfb_im2col_f16")
        self.launches += 1
        return out, Ho, Wo

    def _pool(self, a, NI, H, W, C, k, s, p, is_max):
        Ho, Wo = (H + 2 * p[0] - k[0]) // s[0] + 1, (W + 2 * p[1] - k[1]) // s[1] + 1
        out = torch.empty(NI * Ho * Wo, C, device=a.device, dtype=torch.float16)
        L.check(L.lib().dfb_pool2d_f16(L.ptr(a), L.ptr(out), NI, H, W, C, k[0], k[1], s[0], s[1], p[0], p[1],
                                       1 if is_max else 0, L.cur_stream()), "dfb_pool2d_f16")
        self.launches += 1
        return out, Ho, Wo

    # -------------------------------------------------------------------------------- video
    @torch.no_grad()
    def encode_video(self, video, normalize=False, train=False, pool=True):
        """video [B,T,3,H,W] in [0,1] -> [B,T,512] (pool=False) or [B,512] (MaxPool1d(16) over T)."""
        if video.device.type != "cuda":
            raise RuntimeError("CAVPInferenceB200 runs on a CUDA (sm_100a) device only -- there is no CPU path")
        P = self._pack_all(video.device)
        B, T, Cin, H, W = video.shape
        with torch.cuda.device(video.device):
            x = F.pad(video.permute(0, 1, 3, 4, 2).reshape(B * T, H, W, Cin), (0, 8 - Cin)).half().contiguous()
            col, H, W = self._im2col(x, B * T, H, W, 8, 7, 7, 2, 3, 448)         # (1,7,7)/2 stem
            x = self._gemm(col, P["stem"], RELU)
            x, H, W = self._pool(x, B * T, H, W, 64, (3, 3), (2, 2), (1, 1), True)   # MaxPool3d (1,3,3)/2
            C = 64
            for li in range(1, 5):
                for bi, blk in enumerate(getattr(self.video_encoder, f"layer{li}")):
                    key = f"l{li}.{bi}."
                    s = blk.conv2.stride[1]
                    h = self._conv(x, (B, T, H, W, C), P[key + "conv1"], blk.conv1.k, RELU)
                    planes = h.shape[1]
                    if s == 1:
                        h = self._conv(h, (B, T, H, W, planes), P[key + "conv2"], (1, 3, 3), RELU)
                        H2, W2 = H, W
                    else:
                        col, H2, W2 = self._im2col(h, B * T, H, W, planes, 3, 3, s, 1, 9 * planes)
                        h = self._gemm(col, P[key + "conv2"], RELU)
                    if hasattr(blk, "downsample"):
                        if s == 1:
                            idt = self._gemm(x, P[key + "downsample"], NONE)
                        else:
                            sub, _, _ = self._im2col(x, B * T, H, W, C, 1, 1, s, 0, C)
                            idt = self._gemm(sub, P[key + "downsample"], NONE)
                    else:
                        idt = x
                    x = self._gemm(h, P[key + "conv3"], RELU, residual=idt)   # relu(conv3 + identity)
                    H, W, C = H2, W2, x.shape[1]
            x, _, _ = self._pool(x, B * T, H, W, C, (H, W), (H, W), (0, 0), False)   # AdaptiveAvgPool2d((1,1))
            feat = self._gemm(x, P["vproj"], NONE, out_dtype=torch.float32).view(B, T, -1)
        if pool:
            feat = self.video_pool(feat.permute(0, 2, 1)).squeeze(2)
        if normalize:
            feat = F.normalize(feat, dim=-1)
        return feat

    # --------------------------------------------------------------------------------- spec
    @torch.no_grad()
    def encode_spec(self, spec, normalize=False, pool=True):
        """spec [B,128,T] -> [B,T/16,512] (pool=False) or [B,512]."""
        if spec.device.type != "cuda":
            raise RuntimeError("CAVPInferenceB200 runs on a CUDA (sm_100a) device only -- there is no CPU path")
        P = self._pack_all(spec.device)
        B, mel, T = spec.shape
        with torch.cuda.device(spec.device):
            s0, t0 = P["spec_bn"]
            x = spec.permute(0, 2, 1).float() * s0 + t0                       # bn over mel bins (:1521-1523)
            x = F.pad(x.reshape(B, T, mel, 1), (0, 7)).half().contiguous()    # [B,T,mel,8]
            H, W, C = T, mel, 8
            pools = [(2, 2), (2, 2), (2, 2), (2, 2), (1, 2), (1, 1)]
            for i in range(1, 7):
                if i == 1:
                    col, _, _ = self._im2col(x, B, H, W, 8, 3, 3, 1, 1, 128)
                    x = self._gemm(col, P["cb1.1"], RELU)
                else:
                    x = self._conv(x, (B, 1, H, W, C), P[f"cb{i}.1"], (1, 3, 3), RELU)
                C = x.shape[1]
                x = self._conv(x, (B, 1, H, W, C), P[f"cb{i}.2"], (1, 3, 3), RELU)
                if pools[i - 1] != (1, 1):
                    x, H, W = self._pool(x, B, H, W, C, pools[i - 1], pools[i - 1], (0, 0), False)
            x, _, _ = self._pool(x, B, H, W, C, (1, W), (1, W), (0, 0), False)   # mean over mel -> [B*H, C]
            mx, _, _ = self._pool(x, B, H, 1, C, (3, 1), (1, 1), (1, 0), True)    # max_pool1d(3,1,1) over time
            av, _, _ = self._pool(x, B, H, 1, C, (3, 1), (1, 1), (1, 0), False)   # avg_pool1d(3,1,1), pad counted
            x = (mx.float() + av.float()).half()
            x = self._gemm(x, P["fc1"], RELU)
            x = self._gemm(x, P["fc1"], RELU)                                  # fc1 twice: reference quirk
            feat = self._gemm(x, P["final"], NONE, out_dtype=torch.float32).view(B, H, -1)
        if pool:
            feat = self.spec_pool(feat.permute(0, 2, 1)).squeeze(2)
        if normalize:
            feat = F.normalize(feat, dim=-1)
        return feat

    def forward(self, video, spec, output_dict=True):
        v = self.encode_video(video, normalize=True)
        s = self.encode_spec(spec, normalize=True)
        if output_dict:
            return {"video_features": v, "spec_features": s, "logit_scale": self.logit_scale.exp()}
        return v, s, self.logit_scale.exp()
