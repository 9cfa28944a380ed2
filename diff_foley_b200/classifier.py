"""Double-guidance alignment classifier: same parameter names / call signature as the
reference's `Alignment_Classifier_Double_Guidance` + `Classifier_Backbone`
(diff_foley/modules/double_guidance/alignment_classifier.py:72-295, alignment_backbone.py:417-686).

STATUS -- the one part of the hot path that is NOT on hand-written kernels yet.  Classifier guidance
needs d/dx of log p(x_t, t, video_feat) through this half-UNet every step (ddim.py:333-341), i.e.
forward AND backward of conv / GroupNorm / attention.  Backward kernels are the "next" item N2 of the
scope table; until they exist the classifier forward/backward runs on torch autograd (cuDNN / cuBLAS
library kernels on the GPU -- never the CPU, never the oracle).  It is 2.4 % of the per-step FLOPs
(2 x 2.87 GFLOP vs 355.7).  What is native already: the guided update itself
(`dfb_ddim_step(..., grad, grad_coef)`, ddim.py:377-395) and the UNet the gradient is added to.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .unet import _Box, _Holder, _res, _st


def _timestep_embedding(t, dim, max_period=10000):
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32, device=t.device) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


class ClassifierBackboneB200(nn.Module):
    """Half UNet -> GN/SiLU/conv -> global average pool -> Linear -> sigmoid
    (alignment_backbone.py:417-686)."""

    def __init__(self, image_size, in_channels, model_channels, out_channels, num_res_blocks,
                 attention_resolutions, dropout=0, channel_mult=(1, 2, 4, 8), conv_resample=True, dims=2,
                 num_classes=None, use_checkpoint=False, use_fp16=False, num_heads=-1, num_head_channels=-1,
                 num_heads_upsample=-1, use_scale_shift_norm=False, resblock_updown=False,
                 use_new_attention_order=False, use_spatial_transformer=False, transformer_depth=1,
                 context_dim=None, n_embed=None, legacy=True):
        super().__init__()
        if (dims != 2 or not conv_resample or num_classes is not None or use_scale_shift_norm or resblock_updown
                or not use_spatial_transformer or transformer_depth != 1 or num_heads == -1 or dropout):
            raise NotImplementedError("ClassifierBackboneB200 covers the Diff-Foley inference configuration only")
        self.model_channels, self.num_heads = model_channels, num_heads
        mc, td = model_channels, 4 * model_channels
        self.time_embed = _Box()
        self.time_embed.put("0", _Holder((td, mc)))
        self.time_embed.put("2", _Holder((td, td)))
        self.input_blocks = _Box()
        self.input_blocks.put(0, _Box()).put("0", _Holder((mc, in_channels, 3, 3)))
        self._layout = [[("stem", "input_blocks.0.0")]]
        ch, ds, n_in = mc, 1, 1
        for level, mult in enumerate(channel_mult):
            for _ in range(num_res_blocks):
                blk = self.input_blocks.put(n_in, _Box())
                blk.put("0", _res(ch, mult * mc, td))
                lay = [("res", f"input_blocks.{n_in}.0")]
                ch = mult * mc
                if ds in attention_resolutions:
                    blk.put("1", _st(ch, context_dim))
                    lay.append(("st", f"input_blocks.{n_in}.1"))
                self._layout.append(lay)
                n_in += 1
            if level != len(channel_mult) - 1:
                blk = self.input_blocks.put(n_in, _Box())
                blk.put("0", _Box()).put("op", _Holder((ch, ch, 3, 3)))
                self._layout.append([("down", f"input_blocks.{n_in}.0.op")])
                ds *= 2
                n_in += 1
        self.middle_block = _Box()
        self.middle_block.put("0", _res(ch, ch, td))
        self.middle_block.put("1", _st(ch, context_dim))
        self.middle_block.put("2", _res(ch, ch, td))
        self._layout.append([("res", "middle_block.0"), ("st", "middle_block.1"), ("res", "middle_block.2")])
        last = model_channels * channel_mult[-1]
        self.out = _Box()
        self.out.put("0", _Holder((ch,), kind="norm"))
        self.out.put("2", _Holder((last // 2, last, 3, 3), kind="zero"))
        self.classifier = _Holder((out_channels, last // 2))

    # ---- functional building blocks on this module's own parameters (autograd-capable)
    def _p(self, path):
        return self.get_submodule(path)

    def _res(self, path, x, emb):
        m = self._p(path)
        h = F.conv2d(F.silu(F.group_norm(x, 32, m.in_layers[0].weight, m.in_layers[0].bias, 1e-5)),
                     m.in_layers[2].weight, m.in_layers[2].bias, padding=1)
        h = h + F.linear(F.silu(emb), m.emb_layers[1].weight, m.emb_layers[1].bias)[:, :, None, None]
        h = F.conv2d(F.silu(F.group_norm(h, 32, m.out_layers[0].weight, m.out_layers[0].bias, 1e-5)),
                     m.out_layers[3].weight, m.out_layers[3].bias, padding=1)
        if hasattr(m, "skip_connection"):
            x = F.conv2d(x, m.skip_connection.weight, m.skip_connection.bias)
        return x + h

    def _attn(self, a, x, ctx):
        heads = self.num_heads
        q, k, v = F.linear(x, a.to_q.weight), F.linear(ctx, a.to_k.weight), F.linear(ctx, a.to_v.weight)
        b, n, c = q.shape
        d = c // heads
        sp = lambda t: t.reshape(b, t.shape[1], heads, d).permute(0, 2, 1, 3)
        q, k, v = sp(q), sp(k), sp(v)
        att = (torch.einsum("bhid,bhjd->bhij", q, k) * d ** -0.5).softmax(dim=-1)
        o = torch.einsum("bhij,bhjd->bhid", att, v).permute(0, 2, 1, 3).reshape(b, n, c)
        return F.linear(o, a.to_out[0].weight, a.to_out[0].bias)

    def _st(self, path, x, ctx):
        m = self._p(path)
        b, c, h, w = x.shape
        x_in = x
        x = F.conv2d(F.group_norm(x, 32, m.norm.weight, m.norm.bias, 1e-6), m.proj_in.weight, m.proj_in.bias)
        x = x.permute(0, 2, 3, 1).reshape(b, h * w, c)
        t = m.transformer_blocks[0]
        ln = lambda n, v: F.layer_norm(v, (c,), n.weight, n.bias, 1e-5)
        y = ln(t.norm1, x)
        x = self._attn(t.attn1, y, y) + x
        x = self._attn(t.attn2, ln(t.norm2, x), ctx) + x
        val, gate = F.linear(ln(t.norm3, x), t.ff.net[0].proj.weight, t.ff.net[0].proj.bias).chunk(2, dim=-1)
        x = F.linear(val * F.gelu(gate), t.ff.net[2].weight, t.ff.net[2].bias) + x
        x = x.reshape(b, h, w, c).permute(0, 3, 1, 2)
        return F.conv2d(x, m.proj_out.weight, m.proj_out.bias) + x_in

    def forward(self, x, timesteps=None, context=None, y=None, **kwargs):
        emb = _timestep_embedding(timesteps, self.model_channels)
        emb = F.linear(emb, self.time_embed[0].weight, self.time_embed[0].bias)
        emb = F.linear(F.silu(emb), self.time_embed[2].weight, self.time_embed[2].bias)
        h = x.float()
        ctx = context.float()
        for blk in self._layout:
            for kind, path in blk:
                if kind == "stem":
                    m = self._p(path)
                    h = F.conv2d(h, m.weight, m.bias, padding=1)
                elif kind == "res":
                    h = self._res(path, h, emb)
                elif kind == "st":
                    h = self._st(path, h, ctx)
                else:
                    m = self._p(path)
                    h = F.conv2d(h, m.weight, m.bias, stride=2, padding=1)
        h = F.conv2d(F.silu(F.group_norm(h, 32, self.out[0].weight, self.out[0].bias, 1e-5)),
                     self.out[2].weight, self.out[2].bias, padding=1)
        h = h.mean(dim=(2, 3))
        return torch.sigmoid(F.linear(h, self.classifier.weight, self.classifier.bias))


# Double_Guidance_Classifier.yaml (inference/config): the 11.45 M-parameter half-UNet
DIFF_FOLEY_CLASSIFIER_PARAMS = dict(image_size=32, in_channels=4, out_channels=1, model_channels=128,
                                    attention_resolutions=[2, 4], num_res_blocks=1, channel_mult=[1, 2, 2],
                                    num_heads=8, use_spatial_transformer=True, transformer_depth=1, context_dim=512,
                                    use_checkpoint=True, legacy=False)


class AlignmentClassifierDoubleGuidanceB200(nn.Module):
    """`.model` / `.cond_model` like the reference wrapper; in inference the raw (un-embedded) CAVP
    features go straight to the backbone (alignment_classifier.py:269-271)."""

    def __init__(self, classifier_params=None, cond_stage_params=None, scale_factor=0.18215, **ignored):
        super().__init__()
        from .ldm import VideoFeatEncoderPosembed
        self.model = ClassifierBackboneB200(**(DIFF_FOLEY_CLASSIFIER_PARAMS if classifier_params is None else classifier_params))
        cs = dict(origin_dim=512, embed_dim=512, seq_len=40) if cond_stage_params is None else cond_stage_params
        self.cond_model = VideoFeatEncoderPosembed(**cs)
        self.register_buffer("scale_factor", torch.tensor(scale_factor))

    def forward(self, spec_noisy, video_feat, t):
        return self.model(spec_noisy, context=video_feat, timesteps=t)

    @staticmethod
    def backend_description():
        return "torch autograd on the GPU (cuDNN / cuBLAS library kernels) -- forward + backward, see classifier.py"
