"""Per-launch event-timed profile of one full-size UNet forward (dfb_unet_profile), grouped by shape."""
import collections
import sys

sys.path.insert(0, ".")
import torch

from bench import FULL
from diff_foley_b200.unet import UNetModelB200
from diff_foley_b200.weights import randomize_parameters_

b_eff = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda", 0)
unet = UNetModelB200(**FULL, max_batch=max(b_eff, 2)).to(dev)
randomize_parameters_(unet, 7)
g = torch.Generator().manual_seed(0)
x = torch.randn(b_eff, 4, 16, 64, generator=g).to(dev)
ctx = torch.randn(b_eff, 32, 768, generator=g).to(dev)
t = torch.full((b_eff,), 961, device=dev, dtype=torch.long)
prof = unet.profile(x, t, ctx, iters=10)
agg = collections.OrderedDict()
for p in prof:
    key = (p["kind"], p["M"], p["N"], p["K"], p["splits"], p["ctas"])
    a = agg.setdefault(key, [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += p["ms"]; a[2] += p["flops"]; a[3] += p["bytes"]
tot = sum(p["ms"] for p in prof)
print(f"b_eff={b_eff} launches={len(prof)} total={tot:.3f} ms")
print(f"{'kind':14s} {'M':>6s} {'N':>6s} {'K':>6s} {'spl':>3s} {'ctas':>5s} {'n':>3s} {'ms':>8s} {'us/launch':>9s} {'TF/s':>7s} {'GB/s':>7s}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    us = a[1] / a[0] * 1e3
    print(f"{k[0]:14s} {k[1]:6d} {k[2]:6d} {k[3]:6d} {k[4]:3d} {k[5]:5d} {a[0]:3d} {a[1]:8.3f} {us:9.1f} "
          f"{a[2] / a[1] / 1e9:7.1f} {a[3] / a[1] / 1e6:7.0f}")
