"""One eager UNet forward (B_eff = 2, full-size model) bracketed by cudaProfilerStart/Stop, for
    ncu --profile-from-start off ... python tools/profile_step.py [b_eff]
(the recipe in /opt/skills/guides/B200_PROFILING.md).  Numbers printed under ncu are not bench values."""
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
ctx[: b_eff // 2] = 0
t = torch.full((b_eff,), 961, device=dev, dtype=torch.long)
for _ in range(2):
    unet(x, t, context=ctx)
torch.cuda.synchronize()
torch.cuda.profiler.start()
unet(x, t, context=ctx)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("launches", unet.last_launch_count())
