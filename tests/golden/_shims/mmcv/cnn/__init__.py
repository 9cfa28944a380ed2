import torch.nn as nn


def build_activation_layer(cfg):
    assert cfg["type"] == "ReLU"
    return nn.ReLU(inplace=cfg.get("inplace", False))


class ConvModule(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 bias="auto", conv_cfg=None, norm_cfg=None, act_cfg=dict(type="ReLU"), inplace=True, **kw):
        super().__init__()
        assert conv_cfg is None or conv_cfg["type"] == "Conv3d"
        with_norm = norm_cfg is not None
        if bias == "auto":
            bias = not with_norm
        self.conv = nn.Conv3d(in_channels, out_channels, kernel_size, stride=stride, padding=padding,
                              dilation=dilation, groups=groups, bias=bias)
        self.with_norm, self.with_activation = with_norm, act_cfg is not None
        if with_norm:
            assert norm_cfg["type"] == "BN3d"
            self.bn = nn.BatchNorm3d(out_channels)
        if self.with_activation:
            self.activate = build_activation_layer(act_cfg)

    @property
    def norm(self):
        return self.bn

    def forward(self, x):
        x = self.conv(x)
        if self.with_norm:
            x = self.bn(x)
        if self.with_activation:
            x = self.activate(x)
        return x


class NonLocal3d(nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("NonLocal3d is not used by the CAVP inference config")


def kaiming_init(module, **kw):
    if hasattr(module, "weight") and module.weight is not None:
        nn.init.kaiming_normal_(module.weight, mode="fan_out", nonlinearity="relu")
    if hasattr(module, "bias") and module.bias is not None:
        nn.init.constant_(module.bias, 0)


def constant_init(module, val, bias=0):
    if hasattr(module, "weight") and module.weight is not None:
        nn.init.constant_(module.weight, val)
    if hasattr(module, "bias") and module.bias is not None:
        nn.init.constant_(module.bias, bias)
