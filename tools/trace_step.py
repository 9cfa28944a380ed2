"""In-kernel timeline of one graph-replayed UNet forward (dfb_unet_trace): per launch, when the first
CTA entered, when the programmatic-dependent-launch wait released, the GEMM phases, and the exit --
all from %globaltimer inside the kernels, so launch gaps and in-kernel phases are separated.
    python tools/trace_step.py [b_eff] > profiles/rN_trace_unet_fwd.log"""
import sys

sys.path.insert(0, ".")
import numpy as np
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
prof = unet.profile(x, t, ctx, iters=2)
tr = unet.trace(x, t, ctx)
n = len(prof)
assert tr.shape[0] == n
ok = tr[:, 0, 0] >= 0
t0 = tr[ok, 0, 0].min()
print(f"# b_eff={b_eff}  {n} launches, traced {int(ok.sum())}; forward = {(tr[ok, 7, 1].max() - t0) / 1e3:.1f} us (first entry -> last exit)")
print("# times in us. entry = first CTA entry since forward start; wait = PDL wait released - entry (first CTA);"
      " body = last exit - first wait release; gap = this first wait release - previous last exit;"
      " igemm phases relative to the wait release: ld = first stage landed, mma = last MMA issued, acc = accumulator complete,"
      " epi = epilogue done (last CTA), red = split-K reduction done (last CTA)")
print(f"{'#':>3} {'kind':14s} {'M':>5} {'N':>6} {'K':>6} {'sp':>2} {'ctas':>4} | {'entry':>8} {'wait':>6} {'body':>6} {'gap':>6} | {'ld':>5} {'mma':>5} {'acc':>5} {'epi':>5} {'red':>5} {'exit':>5}")
prev_exit = None
agg = {}
for i in range(n):
    p = prof[i]
    if not ok[i]:
        print(f"{i:3d} {p['kind']:14s} (not instrumented)")
        prev_exit = None
        continue
    e0, w0, w1 = tr[i, 0, 0], tr[i, 1, 0], tr[i, 1, 1]
    x1 = tr[i, 7, 1]
    rel = lambda k, j: (tr[i, k, j] - w0) / 1e3 if tr[i, k, 0] >= 0 else float("nan")
    gap = (w0 - prev_exit) / 1e3 if prev_exit is not None else float("nan")
    body = (x1 - w0) / 1e3
    print(f"{i:3d} {p['kind']:14s} {p['M']:5d} {p['N']:6d} {p['K']:6d} {p['splits']:2d} {p['ctas']:4d} | {(e0 - t0) / 1e3:8.1f} {(w0 - e0) / 1e3:6.1f} {body:6.1f} {gap:6.1f} |"
          f" {rel(2, 0):5.1f} {rel(3, 1):5.1f} {rel(4, 1):5.1f} {rel(5, 1):5.1f} {rel(6, 1):5.1f} {rel(7, 1):5.1f}"
          + ("" if tr[i, 8, 0] < 0 else f" | p1 {rel(12, 0):5.1f} {rel(13, 0):5.1f} cp {rel(15, 0):5.1f} {rel(15, 1):5.1f} | p2 {rel(8, 0):5.1f} {rel(8, 1):5.1f} sum {rel(14, 0):5.1f} b {rel(9, 0):5.1f} {rel(10, 0):5.1f} {rel(11, 0):5.1f}"))
    a = agg.setdefault(p["kind"], [0, 0.0, 0.0])
    a[0] += 1; a[1] += body; a[2] += 0.0 if gap != gap else gap
    prev_exit = x1
print("# totals by kind: launches, sum body us, sum gap us")
for k, a in sorted(agg.items()):
    print(f"# {k:14s} {a[0]:4d} {a[1]:9.1f} {a[2]:9.1f}")
