"""3x3 conv shapes of the B_eff = 16 / 2 plans in isolation (graph-replayed, event-timed)."""
import sys
sys.path.insert(0, ".")
import torch
from diff_foley_b200 import _lib as L
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
splits = int(sys.argv[2]) if len(sys.argv) > 2 else 0
dev = "cuda"; lib = L.lib()
for (B, H, W, C, N) in [(16, 16, 64, 320, 320), (16, 16, 64, 640, 320), (16, 16, 64, 640, 640), (16, 8, 32, 640, 640), (16, 8, 32, 1280, 640),
                        (16, 8, 32, 1280, 1280), (16, 4, 16, 1280, 1280), (2, 16, 64, 320, 320), (2, 16, 64, 640, 640), (2, 8, 32, 640, 640)]:
    a = torch.randn(B, H, W, C, device=dev).half()
    w = (torch.randn(N, 9 * C, device=dev) / (9 * C) ** 0.5).half()
    bias = torch.randn(N, device=dev)
    out = torch.empty(B * H * W, N, device=dev)
    fn = lambda: L.check(lib.dfb_conv3x3(L.ptr(a), L.ptr(w), B, H, W, C, N, L.ptr(bias), None, None, 0, L.ptr(out), None, splits, L.cur_stream()), "conv")
    fn(); torch.cuda.synchronize()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(reps): fn()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): g.replay()
        e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (5 * reps)
    M = B * H * W
    print(f"M={M:6d} N={N:5d} K={9 * C:6d}: {us:8.2f} us  {2.0 * M * N * 9 * C / us / 1e6:7.1f} TFLOP/s")
