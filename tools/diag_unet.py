"""Block-by-block comparison of the CUDA engine against the fp32 oracle (run on the GPU box)."""
import os
import sys

os.environ["DFB_DEBUG_TAPS"] = "1"
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import numpy as np
import torch

from diff_foley_b200.unet import UNetModelB200
from helpers import unet_kwargs
from oracle import unet_oracle

which = sys.argv[1] if len(sys.argv) > 1 else "small"
cfg = unet_oracle.small_unet_cfg() if which == "small" else unet_oracle.DIFF_FOLEY_UNET
g = np.load(f"tests/golden/unet_{which}.npz")
sd = unet_oracle.seeded_state_dict(cfg, int(g["seed"]))
x, t, ctx = (torch.from_numpy(g[k]) for k in ("x", "t", "ctx"))
taps = {}
ref = unet_oracle.unet_forward(sd, cfg, x, t, ctx, taps)
m = UNetModelB200(**unet_kwargs(cfg))
m.load_state_dict(sd)
m = m.cuda()
out = m(x.cuda(), t.cuda(), context=ctx.cuda())
torch.cuda.synchronize()
got = m.debug_taps(x.shape[0])
rel = lambda a, b: float((a.double().cpu() - b.double()).norm() / b.double().norm())
for k, v in got.items():
    print(f"{k:20s} {tuple(v.shape)}  rel-L2 = {rel(v, taps[k]):.3e}")
print("eps rel-L2 =", rel(out, ref), " vs golden", rel(out, torch.from_numpy(g["eps"])))
