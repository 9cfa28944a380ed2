"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by the product path (diff_foley_b200/).

CPU restatement of the first-stage decode that turns a latent into the mel-spectrogram on which the
north-star tolerance is stated (row a19 / N1: measurement path; the product does not ship a VAE yet):
  diff_foley/models/diffusion/ddpm.py:739-797            decode_first_stage: z / scale_factor
  diff_foley/models/autoencoder.py:330-333               AutoencoderKL.decode: post_quant_conv -> Decoder
  diff_foley/modules/stage1_autoencoder/model.py:557-663 Decoder.forward (ResnetBlock :177-242,
                                                         AttnBlock :245-300, Upsample :137-152, swish :128)
Config: inference/config/Stage2_LDM.yaml:38-57 (ch 128, ch_mult 1,2,4,4, 2 res blocks, no attn levels).
Pinned by tests/golden/make_golden.py (vae_decode.npz, generated with the reference's AutoencoderKL).
"""
import math
from collections import OrderedDict

import torch
import torch.nn.functional as F

VAE_CFG = dict(ch=128, out_ch=3, ch_mult=(1, 2, 4, 4), num_res_blocks=2, z_channels=4, embed_dim=4)
SCALE_FACTOR = 0.18215


def decoder_param_shapes(cfg=VAE_CFG):
    s = OrderedDict()

    def conv(p, n, c, k):
        s[p + ".weight"] = (n, c, k, k)
        s[p + ".bias"] = (n,)

    def norm(p, c):
        s[p + ".weight"] = (c,)
        s[p + ".bias"] = (c,)

    def res(p, cin, cout):
        norm(p + ".norm1", cin); conv(p + ".conv1", cout, cin, 3)
        norm(p + ".norm2", cout); conv(p + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(p + ".nin_shortcut", cout, cin, 1)

    ch, mults = cfg["ch"], cfg["ch_mult"]
    block_in = ch * mults[-1]
    conv("post_quant_conv", cfg["z_channels"], cfg["embed_dim"], 1)
    conv("decoder.conv_in", block_in, cfg["z_channels"], 3)
    res("decoder.mid.block_1", block_in, block_in)
    norm("decoder.mid.attn_1.norm", block_in)
    for n in ("q", "k", "v", "proj_out"):
        conv("decoder.mid.attn_1." + n, block_in, block_in, 1)
    res("decoder.mid.block_2", block_in, block_in)
    for lvl in reversed(range(len(mults))):
        block_out = ch * mults[lvl]
        for i in range(cfg["num_res_blocks"] + 1):
            res(f"decoder.up.{lvl}.block.{i}", block_in, block_out)
            block_in = block_out
        if lvl != 0:
            conv(f"decoder.up.{lvl}.upsample.conv", block_in, block_in, 3)
    norm("decoder.norm_out", block_in)
    conv("decoder.conv_out", cfg["out_ch"], block_in, 3)
    return s


def seeded_state_dict(seed=0, cfg=VAE_CFG):
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    for name, shp in decoder_param_shapes(cfg).items():
        if len(shp) == 1 and "norm" in name:
            sd[name] = 1.0 + 0.1 * torch.randn(shp, generator=g) if name.endswith("weight") else 0.1 * torch.randn(shp, generator=g)
        elif len(shp) == 4:
            b = 1.0 / math.sqrt(shp[1] * shp[2] * shp[3])
            sd[name] = (torch.rand(shp, generator=g) * 2 - 1) * b
            sd[name[:-6] + "bias"] = None  # filled right below, keeps registration order
        else:
            w = sd[name[:-4] + "weight"]
            b = 1.0 / math.sqrt(w.shape[1] * w.shape[2] * w.shape[3])
            sd[name] = (torch.rand(shp, generator=g) * 2 - 1) * b
    return sd


def _swish(x):
    return x * torch.sigmoid(x)


def _gn(sd, p, x):
    return F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], 1e-6)


def _res(sd, p, x):
    h = F.conv2d(_swish(_gn(sd, p + ".norm1", x)), sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], padding=1)
    h = F.conv2d(_swish(_gn(sd, p + ".norm2", h)), sd[p + ".conv2.weight"], sd[p + ".conv2.bias"], padding=1)
    if (p + ".nin_shortcut.weight") in sd:
        x = F.conv2d(x, sd[p + ".nin_shortcut.weight"], sd[p + ".nin_shortcut.bias"])
    return x + h


def _attn(sd, p, x):
    h = _gn(sd, p + ".norm", x)
    q, k, v = (F.conv2d(h, sd[f"{p}.{n}.weight"], sd[f"{p}.{n}.bias"]) for n in ("q", "k", "v"))
    b, c, hh, ww = q.shape
    q = q.reshape(b, c, hh * ww).permute(0, 2, 1)
    k = k.reshape(b, c, hh * ww)
    w_ = torch.softmax(torch.bmm(q, k) * (int(c) ** -0.5), dim=2)
    v = v.reshape(b, c, hh * ww)
    h = torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, hh, ww)
    return x + F.conv2d(h, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])


@torch.no_grad()
def decode_first_stage(sd, z, cfg=VAE_CFG, scale_factor=SCALE_FACTOR):
    """latent [B,4,16,64] -> image [B,3,128,512]; channel 0 is the mel-spectrogram."""
    z = (1.0 / scale_factor) * z
    z = F.conv2d(z, sd["post_quant_conv.weight"], sd["post_quant_conv.bias"])
    h = F.conv2d(z, sd["decoder.conv_in.weight"], sd["decoder.conv_in.bias"], padding=1)
    h = _res(sd, "decoder.mid.block_1", h)
    h = _attn(sd, "decoder.mid.attn_1", h)
    h = _res(sd, "decoder.mid.block_2", h)
    for lvl in reversed(range(len(cfg["ch_mult"]))):
        for i in range(cfg["num_res_blocks"] + 1):
            h = _res(sd, f"decoder.up.{lvl}.block.{i}", h)
        if lvl != 0:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = F.conv2d(h, sd[f"decoder.up.{lvl}.upsample.conv.weight"], sd[f"decoder.up.{lvl}.upsample.conv.bias"], padding=1)
    h = _swish(_gn(sd, "decoder.norm_out", h))
    return F.conv2d(h, sd["decoder.conv_out.weight"], sd["decoder.conv_out.bias"], padding=1)
