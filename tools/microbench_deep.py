"""Deep-layer (weight-streaming) shapes with cold weights, graph-replayed: conv3x3 at 4x16 and 2x8,
linear 1280x1280 / 5120 at M = 128 / 32."""
import sys
sys.path.insert(0, ".")
import torch
from diff_foley_b200 import _lib as L
dev = "cuda"; lib = L.lib()

def run_graph(fns, iters=5):
    for f in fns: f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for f in fns: f()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (len(fns) * iters)

def conv_fns(B, H, W, C, N, nw, splits=0):
    a = torch.randn(B, H, W, C, device=dev).half()
    ws = [(torch.randn(N, 9 * C, device=dev) / (9 * C) ** 0.5).half() for _ in range(nw)]
    o = torch.empty(B, H, W, N, device=dev)
    def mk(w):
        def f(): L.check(lib.dfb_conv3x3(L.ptr(a), L.ptr(w), B, H, W, C, N, None, None, None, 0, L.ptr(o), None, splits, L.cur_stream()))
        return f
    return [mk(w) for w in ws]

def lin_fns(M, N, K, nw, splits=0):
    a = torch.randn(M, K, device=dev).half()
    ws = [(torch.randn(N, K, device=dev) / K ** 0.5).half() for _ in range(nw)]
    o = torch.empty(M, N, device=dev)
    def mk(w):
        def f(): L.check(lib.dfb_gemm(L.ptr(a), L.ptr(w), M, N, K, None, None, 0, L.ptr(o), None, splits, L.cur_stream()))
        return f
    return [mk(w) for w in ws]

def rep(name, us, mb):
    print(f"{name:46s} {us:8.2f} us   {mb / us * 1e3:8.0f} GB/s of weights", flush=True)

rep("conv3x3 2x4x16  1280->1280 (29.5 MB, cold)", run_graph(conv_fns(2, 4, 16, 1280, 1280, 12)), 29.5)
rep("conv3x3 2x2x8   1280->1280 (29.5 MB, cold)", run_graph(conv_fns(2, 2, 8, 1280, 1280, 12)), 29.5)
rep("conv3x3 2x2x8   2560->1280 (59 MB, cold)", run_graph(conv_fns(2, 2, 8, 2560, 1280, 6)), 59.0)
rep("conv3x3 2x8x32  640->640   (7.4 MB, cold)", run_graph(conv_fns(2, 8, 32, 640, 640, 24)), 7.4)
rep("conv3x3 2x16x64 320->320   (1.8 MB, cold)", run_graph(conv_fns(2, 16, 64, 320, 320, 48)), 1.8)
rep("linear 128x1280x1280  (3.3 MB, cold)", run_graph(lin_fns(128, 1280, 1280, 64)), 3.3)
rep("linear 128x10240x1280 (26 MB, cold)", run_graph(lin_fns(128, 10240, 1280, 12)), 26.2)
rep("linear 128x1280x5120  (13 MB, cold)", run_graph(lin_fns(128, 1280, 5120, 16)), 13.1)
rep("linear 32x1280x1280   (3.3 MB, cold)", run_graph(lin_fns(32, 1280, 1280, 64)), 3.3)
rep("linear 512x640x640    (0.8 MB, cold)", run_graph(lin_fns(512, 640, 640, 64)), 0.8)
rep("linear 2048x320x320   (0.2 MB, cold)", run_graph(lin_fns(2048, 320, 320, 64)), 0.2)
for sp in (1, 2, 4, 8):
    rep(f"conv3x3 2x4x16 1280->1280 splits={sp}", run_graph(conv_fns(2, 4, 16, 1280, 1280, 12, sp)), 29.5)
