"""Microbenchmarks of the igemm kernel in isolation (CUDA-graph replay of back-to-back launches,
event-timed): fixed launch cost, split-K cost, interleaving with small-smem kernels."""
import sys

sys.path.insert(0, ".")
import torch

from diff_foley_b200 import _lib as L

dev = "cuda"
lib = L.lib()


def mk(M, N, K):
    a = torch.randn(M, K, device=dev).half()
    w = (torch.randn(N, K, device=dev) / K ** 0.5).half()
    o32 = torch.empty(M, N, device=dev)
    return a, w, o32


def run_graph(fn, reps=200, iters=5):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * iters)


def gemm_fn(M, N, K, splits=0, residual=False, f16out=False):
    a, w, o = mk(M, N, K)
    res = torch.randn(M, N, device=dev) if residual else None
    o16 = torch.empty(M, N, device=dev, dtype=torch.float16) if f16out else None
    def f():
        L.check(lib.dfb_gemm(L.ptr(a), L.ptr(w), M, N, K, None, L.ptr(res), 0, None if f16out else L.ptr(o),
                             L.ptr(o16), splits, L.cur_stream()))
    return f


def ln_fn(rows, C):
    x = torch.randn(rows, C, device=dev)
    gm = torch.ones(C, device=dev); bt = torch.zeros(C, device=dev)
    o = torch.empty(rows, C, device=dev, dtype=torch.float16)
    def f():
        L.check(lib.dfb_layernorm(L.ptr(x), rows, C, L.ptr(gm), L.ptr(bt), 1e-5, L.ptr(o), L.cur_stream()))
    return f


def report(name, us, flops=None, bytes_=None):
    s = f"{name:58s} {us:8.2f} us"
    if flops: s += f"  {flops / us / 1e6:8.1f} TFLOP/s"
    if bytes_: s += f"  {bytes_ / us / 1e3:8.1f} GB/s"
    print(s, flush=True)


print(torch.cuda.get_device_name(0))
report("tiny gemm 128x128x64", run_graph(gemm_fn(128, 128, 64, 1)))
report("gemm 2048x320x320 (BN64) f32 out", run_graph(gemm_fn(2048, 320, 320, 1)), 2 * 2048 * 320 * 320)
report("gemm 2048x320x320 (BN64) f32 out + residual", run_graph(gemm_fn(2048, 320, 320, 1, residual=True)), 2 * 2048 * 320 * 320)
report("gemm 2048x320x320 (BN64) f16 out", run_graph(gemm_fn(2048, 320, 320, 1, f16out=True)), 2 * 2048 * 320 * 320)
report("gemm 2048x1152x320 f16 out", run_graph(gemm_fn(2048, 1152, 320, 1, f16out=True)), 2 * 2048 * 1152 * 320)
report("gemm 2048x2560x320 f16 out", run_graph(gemm_fn(2048, 2560, 320, 1, f16out=True)), 2 * 2048 * 2560 * 320)
report("gemm 2048x320x1280 f16 out auto-split", run_graph(gemm_fn(2048, 320, 1280, 0, f16out=True)), 2 * 2048 * 320 * 1280)
g1, l1 = gemm_fn(2048, 320, 320, 1), ln_fn(2048, 320)
def alt():
    g1(); l1()
report("alternating gemm 2048x320x320 + layernorm (per pair)", run_graph(alt, 100) , None)
report("layernorm 2048x320 alone", run_graph(l1))
for sp in (1, 2, 4, 7, 14):
    report(f"gemm 128x1280x11520 splits={sp}", run_graph(gemm_fn(128, 1280, 11520, sp), 50), 2 * 128 * 1280 * 11520, 1280 * 11520 * 2)
for sp in (1, 5, 10):
    report(f"gemm 32x1280x1280 splits={sp}", run_graph(gemm_fn(32, 1280, 1280, sp), 100), None, 1280 * 1280 * 2)
report("gemm 128x10240x1280 splits=1", run_graph(gemm_fn(128, 10240, 1280, 1, f16out=True), 50), 2 * 128 * 10240 * 1280, 10240 * 1280 * 2)
report("gemm 8192x1280x1280 splits=1 (compute-bound)", run_graph(gemm_fn(8192, 1280, 1280, 1, f16out=True), 20), 2 * 8192 * 1280 * 1280)
report("gemm 16384x2560x2560 splits=1 (compute-bound)", run_graph(gemm_fn(16384, 2560, 2560, 1, f16out=True), 5), 2 * 16384 * 2560 * 2560)
