"""Times specific GEMM shapes of the UNet plans in isolation (graph-replayed, event-timed; operands L2-warm).
    python tools/microbench_shapes.py [reps]
Use DFB_DEBUG_SKIP=1|2|4 (results wrong) to cost the epilogue phases."""
import sys

sys.path.insert(0, ".")
import torch

from diff_foley_b200 import _lib as L

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
dev = "cuda"
lib = L.lib()
# (M, N, K, act, residual?, out32?, out16?)
SHAPES = [(16384, 2560, 320, 2, 0, 0, 1), (16384, 320, 320, 0, 1, 1, 0), (16384, 1152, 320, 0, 0, 0, 1),
          (4096, 640, 640, 0, 1, 1, 0), (4096, 5120, 640, 2, 0, 0, 1), (1024, 1280, 1280, 0, 1, 1, 0),
          (16384, 320, 1600, 0, 1, 1, 0), (2048, 2560, 320, 2, 0, 0, 1), (2048, 320, 320, 0, 1, 1, 0)]
for (M, N, K, act, res, o32, o16) in SHAPES:
    a = torch.randn(M, K, device=dev).half()
    w = (torch.randn(N, K, device=dev) / K ** 0.5).half()
    bias = torch.randn(N, device=dev)
    No = N // 2 if act == 2 else N
    r = torch.randn(M, No, device=dev) if res else None
    out32 = torch.empty(M, No, device=dev) if o32 else None
    out16 = torch.empty(M, No, device=dev, dtype=torch.float16) if o16 else None

    def fn():
        L.check(lib.dfb_gemm(L.ptr(a), L.ptr(w), M, N, K, L.ptr(bias), L.ptr(r), act, L.ptr(out32), L.ptr(out16), 0,
                             L.cur_stream()), "gemm")
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
    us = e0.elapsed_time(e1) * 1e3 / (5 * reps)
    byt = 2.0 * M * K + 2.0 * N * K + (4.0 * M * No if res else 0) + (4.0 * M * No if o32 else 0) + (2.0 * M * No if o16 else 0)
    print(f"M={M:6d} N={N:5d} K={K:5d} act={act} res={res}: {us:8.2f} us  {2.0 * M * N * K / us / 1e6:7.1f} TFLOP/s  {byt / us / 1e3:7.0f} GB/s")
