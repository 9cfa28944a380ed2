"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by the product path (diff_foley_b200/).

CPU restatement, in plain functional torch (fp32, or fp64 for an error-budget ceiling), of the
reference's UNet denoiser forward for the Diff-Foley inference configuration.  It is pinned against
the reference's own modules by tests/golden/make_golden.py (run in the build container, where
/root/reference is importable) -> tests/golden/*.npz, and checked by tests/test_oracle.py.

Every function cites the reference file:line it restates (luosiallen/Diff-Foley @ 0ba1e8ad):
  diff_foley/modules/diffusionmodules/openai_unetmodel.py   (UNetModel, ResBlock, Up/Downsample)
  diff_foley/modules/diffusionmodules/attention_openai.py   (SpatialTransformer, CrossAttention, GEGLU)
  diff_foley/modules/diffusionmodules/util.py               (timestep_embedding, GroupNorm32)

Only the subset of constructor options the inference YAMLs use is covered
(inference/config/Stage2_LDM.yaml:21-36): dims=2, conv_resample, no scale-shift norm, no
resblock_updown, num_classes=None, use_spatial_transformer with transformer_depth=1, legacy=False.
"""
import math
from collections import OrderedDict

import torch
import torch.nn.functional as F

# inference/config/Stage2_LDM.yaml:21-36 (+ latent geometry ddpm.py:1293, context notebook cell 13)
DIFF_FOLEY_UNET = dict(
    in_channels=4, out_channels=4, model_channels=320, attention_resolutions=(4, 2, 1),
    num_res_blocks=2, channel_mult=(1, 2, 4, 4), num_heads=8, context_dim=768,
    latent_h=16, latent_w=64, context_len=32,
)


def small_unet_cfg(model_channels=64, channel_mult=(1, 2, 4, 4), num_heads=4, context_dim=128,
                   latent_h=16, latent_w=64, context_len=32, num_res_blocks=2,
                   attention_resolutions=(4, 2, 1)):
    """Reduced-width variant of the same architecture for fast CPU tests / small golden files."""
    return dict(in_channels=4, out_channels=4, model_channels=model_channels,
                attention_resolutions=tuple(attention_resolutions), num_res_blocks=num_res_blocks,
                channel_mult=tuple(channel_mult), num_heads=num_heads, context_dim=context_dim,
                latent_h=latent_h, latent_w=latent_w, context_len=context_len)


# ------------------------------------------------------------------------------------ structure
def unet_structure(cfg):
    """Block layout: mirrors the constructor loops at openai_unetmodel.py:513-680.

    Returns (input_blocks, middle_block, output_blocks); each block is a list of layer tuples
    ('res', prefix, cin, cout) | ('st', prefix, C) | ('down', prefix, C) | ('up', prefix, C).
    input_blocks[0] is the stem conv ('stem', prefix, cin, cout).
    """
    mc = cfg["model_channels"]
    attn = set(cfg["attention_resolutions"])
    nres = cfg["num_res_blocks"]
    mults = cfg["channel_mult"]
    inputs = [[("stem", "input_blocks.0.0", cfg["in_channels"], mc)]]
    chans = [mc]
    ch, ds = mc, 1
    for level, mult in enumerate(mults):
        for _ in range(nres):                                   # :525-556
            p = f"input_blocks.{len(inputs)}"
            blk = [("res", p + ".0", ch, mult * mc)]
            ch = mult * mc
            if ds in attn:
                blk.append(("st", p + ".1", ch))
            inputs.append(blk)
            chans.append(ch)
        if level != len(mults) - 1:                             # :557-580
            p = f"input_blocks.{len(inputs)}"
            inputs.append([("down", p + ".0.op", ch)])
            chans.append(ch)
            ds *= 2
    middle = [("res", "middle_block.0", ch, ch), ("st", "middle_block.1", ch),
              ("res", "middle_block.2", ch, ch)]                 # :590-617
    outputs = []
    for level, mult in list(enumerate(mults))[::-1]:            # :620-677
        for i in range(nres + 1):
            ich = chans.pop()
            p = f"output_blocks.{len(outputs)}"
            blk = [("res", p + ".0", ch + ich, mc * mult)]
            ch = mc * mult
            if ds in attn:
                blk.append(("st", p + ".1", ch))
            if level and i == nres:
                blk.append(("up", f"{p}.{len(blk)}.conv", ch))
                ds //= 2
            outputs.append(blk)
    return inputs, middle, outputs


def unet_param_shapes(cfg):
    """state-dict key -> shape, in the reference module's registration order
    (ResBlock: openai_unetmodel.py:201-241; SpatialTransformer: attention_openai.py:226-248;
    BasicTransformerBlock :196-209; CrossAttention :152-168; GEGLU/FeedForward :37-61)."""
    mc, td, cd = cfg["model_channels"], 4 * cfg["model_channels"], cfg["context_dim"]
    s = OrderedDict()

    def lin(p, n, k, bias=True):
        s[p + ".weight"] = (n, k)
        if bias:
            s[p + ".bias"] = (n,)

    def conv(p, n, c, k):
        s[p + ".weight"] = (n, c, k, k)
        s[p + ".bias"] = (n,)

    def norm(p, c):
        s[p + ".weight"] = (c,)
        s[p + ".bias"] = (c,)

    def res(p, cin, cout):
        norm(p + ".in_layers.0", cin)
        conv(p + ".in_layers.2", cout, cin, 3)
        lin(p + ".emb_layers.1", cout, td)
        norm(p + ".out_layers.0", cout)
        conv(p + ".out_layers.3", cout, cout, 3)
        if cin != cout:
            conv(p + ".skip_connection", cout, cin, 1)

    def st(p, c):
        norm(p + ".norm", c)
        conv(p + ".proj_in", c, c, 1)
        t = p + ".transformer_blocks.0"
        for a, kd in (("attn1", c), ("attn2", cd)):
            if a == "attn2":
                pass
            lin(f"{t}.{a}.to_q", c, c, bias=False)
            lin(f"{t}.{a}.to_k", c, kd, bias=False)
            lin(f"{t}.{a}.to_v", c, kd, bias=False)
            lin(f"{t}.{a}.to_out.0", c, c)
            if a == "attn1":
                lin(f"{t}.ff.net.0.proj", 8 * c, c)
                lin(f"{t}.ff.net.2", c, 4 * c)
        norm(t + ".norm1", c)
        norm(t + ".norm2", c)
        norm(t + ".norm3", c)
        conv(p + ".proj_out", c, c, 1)

    lin("time_embed.0", td, mc)
    lin("time_embed.2", td, td)
    inputs, middle, outputs = unet_structure(cfg)
    for blk in inputs + [middle] + outputs:
        for layer in blk:
            kind, p = layer[0], layer[1]
            if kind == "stem":
                conv(p, layer[3], layer[2], 3)
            elif kind == "res":
                res(p, layer[2], layer[3])
            elif kind == "st":
                st(p, layer[2])
            else:
                conv(p, layer[2], layer[2], 3)
    norm("out.0", mc)
    conv("out.2", cfg["out_channels"], mc, 3)
    return s


def seeded_state_dict(cfg, seed=0, dtype=torch.float32):
    """Deterministic random weights for every parameter, INCLUDING the tensors the reference
    zero-initialises (zero_module: openai_unetmodel.py:229-231,685; attention_openai.py:244-248) --
    with those left at zero the UNet output is identically 0 and parity would be vacuous
    (SURVEY F5).  U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weights and biases (the scale of
    torch's default init), norm scales 1 + 0.1 N(0,1), norm shifts 0.1 N(0,1).
    Generated tensor-by-tensor from one torch CPU generator: identical on every host."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    shapes = unet_param_shapes(cfg)
    sd = OrderedDict()
    fan = {}
    for name, shp in shapes.items():
        base = name.rsplit(".", 1)[0]
        is_norm = len(shp) == 1 and (base + ".weight") in shapes and len(shapes[base + ".weight"]) == 1
        if is_norm:
            if name.endswith(".weight"):
                sd[name] = (1.0 + 0.1 * torch.randn(shp, generator=g)).to(dtype)
            else:
                sd[name] = (0.1 * torch.randn(shp, generator=g)).to(dtype)
            continue
        if name.endswith(".weight"):
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            fan[base] = fan_in
        bound = 1.0 / math.sqrt(fan[base])
        sd[name] = ((torch.rand(shp, generator=g) * 2 - 1) * bound).to(dtype)
    return sd


# -------------------------------------------------------------------------------------- forward
def timestep_embedding(t, dim, max_period=10000):
    """util.py:151-171 (repeat_only=False)."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32, device=t.device) / half)
    args = t[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def _gn(sd, p, x, eps):
    # GroupNorm32 computes in fp32 (util.py:214-216); Normalize uses eps 1e-6 (attention_openai.py:76-77)
    return F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], eps)


def resblock(sd, p, x, emb):
    """ResBlock._forward, openai_unetmodel.py:255-275 (no up/down, no scale-shift norm)."""
    h = F.conv2d(F.silu(_gn(sd, p + ".in_layers.0", x, 1e-5)), sd[p + ".in_layers.2.weight"],
                 sd[p + ".in_layers.2.bias"], padding=1)
    emb_out = F.linear(F.silu(emb), sd[p + ".emb_layers.1.weight"], sd[p + ".emb_layers.1.bias"])
    h = h + emb_out[:, :, None, None]
    h = F.conv2d(F.silu(_gn(sd, p + ".out_layers.0", h, 1e-5)), sd[p + ".out_layers.3.weight"],
                 sd[p + ".out_layers.3.bias"], padding=1)
    if (p + ".skip_connection.weight") in sd:
        x = F.conv2d(x, sd[p + ".skip_connection.weight"], sd[p + ".skip_connection.bias"])
    return x + h


def cross_attention(sd, p, x, context, heads):
    """CrossAttention.forward, attention_openai.py:170-193."""
    ctx = x if context is None else context
    q = F.linear(x, sd[p + ".to_q.weight"])
    k = F.linear(ctx, sd[p + ".to_k.weight"])
    v = F.linear(ctx, sd[p + ".to_v.weight"])
    b, n, c = q.shape
    d = c // heads

    def split(t):
        return t.reshape(b, t.shape[1], heads, d).permute(0, 2, 1, 3)

    q, k, v = split(q), split(k), split(v)
    sim = torch.einsum("bhid,bhjd->bhij", q, k) * (d ** -0.5)
    attn = sim.softmax(dim=-1)
    out = torch.einsum("bhij,bhjd->bhid", attn, v).permute(0, 2, 1, 3).reshape(b, n, c)
    return F.linear(out, sd[p + ".to_out.0.weight"], sd[p + ".to_out.0.bias"])


def spatial_transformer(sd, p, x, context, heads):
    """SpatialTransformer.forward :250-261 with one BasicTransformerBlock :211-215, GEGLU :42-44."""
    b, c, h, w = x.shape
    x_in = x
    x = _gn(sd, p + ".norm", x, 1e-6)
    x = F.conv2d(x, sd[p + ".proj_in.weight"], sd[p + ".proj_in.bias"])
    x = x.permute(0, 2, 3, 1).reshape(b, h * w, c)
    t = p + ".transformer_blocks.0"

    def ln(name, v):
        return F.layer_norm(v, (c,), sd[f"{t}.{name}.weight"], sd[f"{t}.{name}.bias"], 1e-5)

    x = cross_attention(sd, t + ".attn1", ln("norm1", x), None, heads) + x
    x = cross_attention(sd, t + ".attn2", ln("norm2", x), context, heads) + x
    proj = F.linear(ln("norm3", x), sd[t + ".ff.net.0.proj.weight"], sd[t + ".ff.net.0.proj.bias"])
    val, gate = proj.chunk(2, dim=-1)
    ff = F.linear(val * F.gelu(gate), sd[t + ".ff.net.2.weight"], sd[t + ".ff.net.2.bias"])
    x = ff + x
    x = x.reshape(b, h, w, c).permute(0, 3, 1, 2)
    x = F.conv2d(x, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])
    return x + x_in


def _run_block(sd, cfg, blk, h, emb, context):
    for layer in blk:
        kind, p = layer[0], layer[1]
        if kind == "stem":
            h = F.conv2d(h, sd[p + ".weight"], sd[p + ".bias"], padding=1)
        elif kind == "res":
            h = resblock(sd, p, h, emb)
        elif kind == "st":
            h = spatial_transformer(sd, p, h, context, cfg["num_heads"])
        elif kind == "down":                                   # Downsample, :134-160
            h = F.conv2d(h, sd[p + ".weight"], sd[p + ".bias"], stride=2, padding=1)
        elif kind == "up":                                     # Upsample, :91-119
            h = F.interpolate(h, scale_factor=2, mode="nearest")
            h = F.conv2d(h, sd[p + ".weight"], sd[p + ".bias"], padding=1)
    return h


@torch.no_grad()
def unet_forward(sd, cfg, x, timesteps, context, taps=None):
    """UNetModel.forward, openai_unetmodel.py:710-742.  `sd` holds tensors of the compute dtype
    (fp32 or fp64) under the reference's state-dict keys.  If `taps` is a dict, the output of every
    block is stored in it (NCHW) for block-level parity checks."""
    dtype = sd["out.2.weight"].dtype
    t_emb = timestep_embedding(timesteps, cfg["model_channels"]).to(dtype)
    emb = F.linear(t_emb, sd["time_embed.0.weight"], sd["time_embed.0.bias"])
    emb = F.linear(F.silu(emb), sd["time_embed.2.weight"], sd["time_embed.2.bias"])
    inputs, middle, outputs = unet_structure(cfg)
    h = x.to(dtype)
    context = context.to(dtype)
    hs = []
    for i, blk in enumerate(inputs):
        h = _run_block(sd, cfg, blk, h, emb, context)
        hs.append(h)
        if taps is not None:
            taps[f"input_blocks.{i}"] = h
    h = _run_block(sd, cfg, middle, h, emb, context)
    if taps is not None:
        taps["middle_block"] = h
    for i, blk in enumerate(outputs):
        h = torch.cat([h, hs.pop()], dim=1)
        h = _run_block(sd, cfg, blk, h, emb, context)
        if taps is not None:
            taps[f"output_blocks.{i}"] = h
    h = F.silu(_gn(sd, "out.0", h, 1e-5))
    return F.conv2d(h, sd["out.2.weight"], sd["out.2.bias"], padding=1)
