"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by the product path (diff_foley_b200/).

CPU restatement of the double-guidance classifier and of its log-likelihood gradient:
  diff_foley/modules/double_guidance/alignment_backbone.py:417-686  Classifier_Backbone
  diff_foley/modules/double_guidance/alignment_classifier.py:269-271 forward (raw features as context)
  diff_foley/models/diffusion/ddim.py:333-341                        cal_classifier_loglikelihood_grad
Building blocks (ResBlock / SpatialTransformer) are the same modules as the UNet's and are shared with
oracle/unet_oracle.py.  Pinned by tests/golden/make_golden.py (classifier_small.npz).
"""
from collections import OrderedDict

import torch
import torch.nn.functional as F

from . import unet_oracle as U

# inference/config/Double_Guidance_Classifier.yaml:36-50
DIFF_FOLEY_CLASSIFIER = dict(in_channels=4, out_channels=1, model_channels=128, attention_resolutions=(2, 4),
                             num_res_blocks=1, channel_mult=(1, 2, 2), num_heads=8, context_dim=512)


def classifier_structure(cfg):
    """input_blocks + middle_block of alignment_backbone.py:521-625 (no output blocks)."""
    full = dict(cfg, latent_h=16, latent_w=64, context_len=32)
    inputs, middle, _ = U.unet_structure(full)
    return inputs, middle


def classifier_param_shapes(cfg):
    full = dict(cfg, out_channels=4, latent_h=16, latent_w=64, context_len=32)
    s = OrderedDict((k, v) for k, v in U.unet_param_shapes(full).items()
                    if not k.startswith("output_blocks.") and not k.startswith("out."))
    last = cfg["model_channels"] * cfg["channel_mult"][-1]
    s["out.0.weight"] = (last,)
    s["out.0.bias"] = (last,)
    s["out.2.weight"] = (last // 2, last, 3, 3)
    s["out.2.bias"] = (last // 2,)
    s["classifier.weight"] = (cfg["out_channels"], last // 2)
    s["classifier.bias"] = (cfg["out_channels"],)
    return s


def seeded_state_dict(cfg, seed=0):
    import math
    g = torch.Generator().manual_seed(seed)
    shapes = classifier_param_shapes(cfg)
    sd, fan = OrderedDict(), {}
    for name, shp in shapes.items():
        base = name.rsplit(".", 1)[0]
        if len(shp) == 1 and len(shapes.get(base + ".weight", (0, 0))) == 1:
            sd[name] = 1.0 + 0.1 * torch.randn(shp, generator=g) if name.endswith(".weight") else 0.1 * torch.randn(shp, generator=g)
            continue
        if name.endswith(".weight"):
            f = 1
            for d in shp[1:]:
                f *= d
            fan[base] = f
        b = 1.0 / math.sqrt(fan[base])
        sd[name] = (torch.rand(shp, generator=g) * 2 - 1) * b
    return sd


def classifier_forward(sd, cfg, x, t, context):
    """Classifier_Backbone.forward, alignment_backbone.py:656-686 -> probabilities [B, out]."""
    full = dict(cfg, latent_h=x.shape[2], latent_w=x.shape[3], context_len=context.shape[1])
    emb = U.timestep_embedding(t, cfg["model_channels"])
    emb = F.linear(emb, sd["time_embed.0.weight"], sd["time_embed.0.bias"])
    emb = F.linear(F.silu(emb), sd["time_embed.2.weight"], sd["time_embed.2.bias"])
    inputs, middle = classifier_structure(cfg)
    h = x
    for blk in inputs:
        h = U._run_block(sd, full, blk, h, emb, context)
    h = U._run_block(sd, full, middle, h, emb, context)
    h = F.silu(F.group_norm(h, 32, sd["out.0.weight"], sd["out.0.bias"], 1e-5))
    h = F.conv2d(h, sd["out.2.weight"], sd["out.2.bias"], padding=1)
    h = h.mean(dim=(2, 3))                                   # AdaptiveAvgPool2d((1,1)) + squeeze
    return torch.sigmoid(F.linear(h, sd["classifier.weight"], sd["classifier.bias"]))


def loglikelihood_grad(sd, cfg, x, t, context, scale):
    """cal_classifier_loglikelihood_grad, ddim.py:333-341."""
    with torch.enable_grad():
        x_in = x.detach().requires_grad_(True)
        logp = torch.log(classifier_forward(sd, cfg, x_in, t, context))
        return torch.autograd.grad(logp.sum(), x_in)[0] * scale
