"""Seeded synthetic parameters for benchmarking / smoke tests (no checkpoint can be downloaded).

The reference zero-initialises every ResBlock out-conv, SpatialTransformer.proj_out and the final
conv (zero_module), so a freshly constructed UNet outputs exactly 0; a benchmark or parity check on
that would be vacuous.  `randomize_parameters_` gives every tensor the scale of torch's
default init instead: U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weights and biases, 1 + 0.1 N(0,1) /
0.1 N(0,1) for norm scales / shifts.
"""
import math

import torch


@torch.no_grad()
def randomize_parameters_(module, seed=0):
    params = dict(module.named_parameters())
    dev = next(iter(params.values())).device
    g = torch.Generator(device=dev).manual_seed(seed)
    for name, p in params.items():
        base = name.rsplit(".", 1)[0]
        w = params.get(base + ".weight")
        if w is not None and w.dim() == 1:  # norm layer
            if name.endswith(".weight"):
                p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g, device=dev))
            else:
                p.copy_(0.1 * torch.randn(p.shape, generator=g, device=dev))
            continue
        ref = w if w is not None else p
        fan_in = 1
        for d in ref.shape[1:]:
            fan_in *= d
        bound = 1.0 / math.sqrt(max(fan_in, 1))
        p.copy_((torch.rand(p.shape, generator=g, device=dev) * 2 - 1) * bound)
    return module
