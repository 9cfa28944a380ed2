"""GroupNorm / LayerNorm kernels in isolation at the UNet's call-site shapes: graph-replayed back-to-back
launches, event-timed, with a max-abs check against torch.  DFB_GN=1 selects the v1 GroupNorm kernel.
    python tools/microbench_norm.py [B_eff] [reps]"""
import sys

sys.path.insert(0, ".")
import torch
import torch.nn.functional as F

from diff_foley_b200 import _lib as L

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
dev = "cuda"
lib = L.lib()
GN = [(320, 1024), (640, 1024), (960, 1024), (320, 256), (640, 256), (960, 256), (1280, 256), (1920, 256),
      (640, 64), (1280, 64), (1920, 64), (2560, 64), (1280, 16), (2560, 16)]


def timeit(fn):
    fn()
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(reps):
                fn()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (5 * reps)


print(f"# groupnorm(+SiLU), B_eff={B}")
tot = 0.0
for C, HW in GN:
    x = torch.randn(B, HW, C, device=dev) * 1.7 + 0.3
    g_ = torch.randn(C, device=dev)
    b_ = torch.randn(C, device=dev)
    out = torch.empty(B, HW, C, device=dev, dtype=torch.float16)
    fn = lambda: L.check(lib.dfb_groupnorm(L.ptr(x), C, None, 0, B, HW, L.ptr(g_), L.ptr(b_), 1e-5, 1, L.ptr(out),
                                           None, L.cur_stream()), "gn")
    us = timeit(fn)
    ref = F.silu(F.group_norm(x.permute(0, 2, 1).reshape(B, C, HW), 32, g_, b_, 1e-5)).reshape(B, C, HW).permute(0, 2, 1)
    err = float((out.float() - ref).abs().max())
    mb = B * HW * C * 6 / 1e6
    tot += us
    print(f"C={C:5d} HW={HW:5d}: {us:7.2f} us  {mb / us * 1e3:7.0f} GB/s   max|err| {err:.2e}")
print(f"# sum {tot:.1f} us")
print(f"# layernorm, B_eff={B}")
for C, rows in [(320, 1024), (640, 256), (1280, 64), (1280, 16)]:
    x = torch.randn(B * rows, C, device=dev)
    g_ = torch.randn(C, device=dev)
    b_ = torch.randn(C, device=dev)
    out = torch.empty(B * rows, C, device=dev, dtype=torch.float16)
    fn = lambda: L.check(lib.dfb_layernorm(L.ptr(x), B * rows, C, L.ptr(g_), L.ptr(b_), 1e-5, L.ptr(out), L.cur_stream()), "ln")
    us = timeit(fn)
    print(f"C={C:5d} rows={B * rows:5d}: {us:7.2f} us  {B * rows * C * 6 / 1e6 / us * 1e3:7.0f} GB/s")
