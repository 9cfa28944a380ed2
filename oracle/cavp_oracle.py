"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by the product path (diff_foley_b200/).

CPU restatement (functional torch fp32) of the CAVP encoders:
  inference/model/cavp_model.py:47-84             encode_video / encode_spec
  inference/model/cavp_modules.py:331-871,1233    ResNet3d / ResNet3dSlowOnly (depth 50, stage blocks
                                                  3,4,6,3; spatial strides 1,2,2,2; inflate 0,0,1,1)
  inference/model/cavp_modules.py:167-328         Bottleneck3d ('pytorch' style: stride on conv2)
  inference/model/cavp_modules.py:1440-1546       ConvBlock / Cnn14 (fc1 applied twice, :1543-1544)
State-dict keys are the reference's (video_encoder.conv1.conv.weight, ...bn.running_mean, ...).
Pinned by tests/golden/make_golden.py (cavp_small.npz) through the mmcv import shim.
"""
import math
from collections import OrderedDict

import torch
import torch.nn.functional as F

STAGES = ((3, 1, 0), (4, 2, 0), (6, 2, 1), (3, 2, 1))  # (blocks, spatial stride, inflate)


def cavp_param_shapes(embed_dim=512):
    s = OrderedDict()

    def convbn(p, cout, cin, k):
        s[p + ".conv.weight"] = (cout, cin) + tuple(k)
        for n, shp in (("weight", (cout,)), ("bias", (cout,)), ("running_mean", (cout,)), ("running_var", (cout,))):
            s[f"{p}.bn.{n}"] = shp

    convbn("video_encoder.conv1", 64, 3, (1, 7, 7))
    inpl = 64
    for i, (blocks, stride, inflate) in enumerate(STAGES):
        planes = 64 * 2 ** i
        for b in range(blocks):
            p = f"video_encoder.layer{i + 1}.{b}"
            st = stride if b == 0 else 1
            convbn(p + ".conv1", planes, inpl, (3, 1, 1) if inflate else (1, 1, 1))
            convbn(p + ".conv2", planes, planes, (1, 3, 3))
            convbn(p + ".conv3", planes * 4, planes, (1, 1, 1))
            if b == 0 and (st != 1 or inpl != planes * 4):
                convbn(p + ".downsample", planes * 4, inpl, (1, 1, 1))
            inpl = planes * 4
    s["video_project_head.weight"] = (embed_dim, 2048)
    s["video_project_head.bias"] = (embed_dim,)
    for n in ("weight", "bias", "running_mean", "running_var"):
        s["spec_encoder.bn." + n] = (128,)
    ch = [1, 64, 128, 256, 512, 1024, 2048]
    for i in range(6):
        p = f"spec_encoder.conv_block{i + 1}"
        s[p + ".conv1.weight"] = (ch[i + 1], ch[i], 3, 3)
        s[p + ".conv2.weight"] = (ch[i + 1], ch[i + 1], 3, 3)
        for j in (1, 2):
            for n in ("weight", "bias", "running_mean", "running_var"):
                s[f"{p}.bn{j}.{n}"] = (ch[i + 1],)
    s["spec_encoder.fc1.weight"] = (2048, 2048)
    s["spec_encoder.fc1.bias"] = (2048,)
    s["spec_encoder.final_project.weight"] = (embed_dim, 2048)
    s["spec_encoder.final_project.bias"] = (embed_dim,)
    return s


def seeded_state_dict(seed=0):
    """He-style conv weights (keeps activations O(1) through 50 layers), randomised BatchNorm affine AND
    running statistics (defaults of 0/1 would hide BN-folding bugs)."""
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    for name, shp in cavp_param_shapes().items():
        if name.endswith("running_var"):
            sd[name] = 0.5 + torch.rand(shp, generator=g)
        elif name.endswith("running_mean"):
            sd[name] = 0.2 * torch.randn(shp, generator=g)
        elif ".bn" in name and name.endswith(".weight"):
            sd[name] = 0.8 + 0.4 * torch.rand(shp, generator=g)
        elif ".bn" in name and name.endswith(".bias"):
            sd[name] = 0.1 * torch.randn(shp, generator=g)
        elif name.endswith(".bias"):
            sd[name] = 0.1 * torch.randn(shp, generator=g)
        else:
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            sd[name] = torch.randn(shp, generator=g) * math.sqrt(1.0 / fan_in)
    return sd


def _bn(sd, p, x):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        False, 0.0, 1e-5)


def _convbn3d(sd, p, x, stride=(1, 1, 1), relu=True):
    w = sd[p + ".conv.weight"]
    x = F.conv3d(x, w, None, stride, tuple(k // 2 for k in w.shape[2:]))
    x = _bn(sd, p + ".bn", x)
    return F.relu(x) if relu else x


@torch.no_grad()
def encode_video(sd, video, normalize=False, pool=True):
    """cavp_model.py:47-65 + ResNet3d.forward cavp_modules.py:837-860."""
    x = video.permute(0, 2, 1, 3, 4)
    x = _convbn3d(sd, "video_encoder.conv1", x, (1, 2, 2))
    x = F.max_pool3d(x, (1, 3, 3), (1, 2, 2), (0, 1, 1))
    for i, (blocks, stride, _) in enumerate(STAGES):
        for b in range(blocks):
            p = f"video_encoder.layer{i + 1}.{b}"
            st = stride if b == 0 else 1
            out = _convbn3d(sd, p + ".conv1", x)
            out = _convbn3d(sd, p + ".conv2", out, (1, st, st))
            out = _convbn3d(sd, p + ".conv3", out, relu=False)
            idt = _convbn3d(sd, p + ".downsample", x, (1, st, st), relu=False) if (p + ".downsample.conv.weight") in sd else x
            x = F.relu(out + idt)
    x = x.mean(dim=(3, 4))                                           # AdaptiveAvgPool2d((1,1)) on 5-D
    x = F.linear(x.permute(0, 2, 1), sd["video_project_head.weight"], sd["video_project_head.bias"])
    if pool:
        x = F.max_pool1d(x.permute(0, 2, 1), 16).squeeze(2)
    return F.normalize(x, dim=-1) if normalize else x


@torch.no_grad()
def encode_spec(sd, spec, normalize=False, pool=True):
    """cavp_model.py:68-84 + Cnn14.forward cavp_modules.py:1516-1546."""
    x = spec.unsqueeze(1).permute(0, 1, 3, 2)                        # B x 1 x T x mel
    x = _bn(sd, "spec_encoder.bn", x.transpose(1, 3)).transpose(1, 3)
    pools = [(2, 2), (2, 2), (2, 2), (2, 2), (1, 2), (1, 1)]
    for i in range(6):
        p = f"spec_encoder.conv_block{i + 1}"
        x = F.relu(_bn(sd, p + ".bn1", F.conv2d(x, sd[p + ".conv1.weight"], None, 1, 1)))
        x = F.relu(_bn(sd, p + ".bn2", F.conv2d(x, sd[p + ".conv2.weight"], None, 1, 1)))
        x = F.avg_pool2d(x, kernel_size=pools[i])
    x = torch.mean(x, dim=3)
    x = F.max_pool1d(x, 3, 1, 1) + F.avg_pool1d(x, 3, 1, 1)
    x = x.transpose(1, 2)
    x = F.relu(F.linear(x, sd["spec_encoder.fc1.weight"], sd["spec_encoder.fc1.bias"]))
    x = F.relu(F.linear(x, sd["spec_encoder.fc1.weight"], sd["spec_encoder.fc1.bias"]))   # sic: fc1 twice
    x = F.linear(x, sd["spec_encoder.final_project.weight"], sd["spec_encoder.final_project.bias"])
    if pool:
        x = F.max_pool1d(x.permute(0, 2, 1), 16).squeeze(2)
    return F.normalize(x, dim=-1) if normalize else x
