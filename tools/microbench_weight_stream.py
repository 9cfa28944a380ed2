"""Weight-streaming convs (M = 32 / 128): L2-warm weights (one copy replayed) vs HBM-cold (12 rotating copies)."""
import sys
sys.path.insert(0, ".")
import torch
from diff_foley_b200 import _lib as L
dev = "cuda"; lib = L.lib()
def bench(B, H, W, C, N, ncopies, reps=12):
    a = torch.randn(B, H, W, C, device=dev).half()
    ws = [(torch.randn(N, 9 * C, device=dev) / (9 * C) ** 0.5).half() for _ in range(ncopies)]
    bias = torch.randn(N, device=dev)
    out = torch.empty(B * H * W, N, device=dev)
    def fn(i): L.check(lib.dfb_conv3x3(L.ptr(a), L.ptr(ws[i % ncopies]), B, H, W, C, N, L.ptr(bias), None, None, 0, L.ptr(out), None, 0, L.cur_stream()), "conv")
    fn(0); torch.cuda.synchronize()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(reps): fn(i)
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): g.replay()
        e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (5 * reps)
for (B, H, W, C, N) in [(2, 2, 8, 1280, 1280), (2, 2, 8, 2560, 1280), (2, 4, 16, 1280, 1280), (2, 4, 16, 2560, 1280), (2, 8, 32, 640, 640)]:
    warm = bench(B, H, W, C, N, 1); cold = bench(B, H, W, C, N, 12)
    mb = N * 9 * C * 2 / 1e6
    print(f"M={B*H*W:4d} N={N} K={9*C:6d} W={mb:5.1f} MB: warm {warm:6.2f} us ({mb/warm:5.2f} TB/s)  cold {cold:6.2f} us ({mb/cold:5.2f} TB/s)")
