"""Cost of the epilogue variants of one small GEMM (2048 x 320 x 320, 80 CTAs): graph of 20 back-to-back launches."""
import os, subprocess, sys, ctypes as C
if len(sys.argv) > 1:
    sys.path.insert(0, ".")
    import torch
    from diff_foley_b200 import _lib as L
    dev = "cuda"; lib = L.lib()
    M, N, K = (int(v) for v in sys.argv[2:5]) if len(sys.argv) > 4 else (2048, 320, 320)
    a = torch.randn(M, K, device=dev).half(); w = (torch.randn(N, K, device=dev) / K ** 0.5).half()
    bias = torch.randn(N, device=dev); res = torch.randn(M, N, device=dev)
    o32 = torch.empty(M, N, device=dev); o16 = torch.empty(M, N, device=dev, dtype=torch.float16)
    stats = torch.empty(M, 8, 2, device=dev); tn = C.c_int(0)
    variants = {
        "fp16 out only": lambda: lib.dfb_gemm(L.ptr(a), L.ptr(w), M, N, K, L.ptr(bias), None, 0, None, L.ptr(o16), 0, L.cur_stream()),
        "fp32 out + residual": lambda: lib.dfb_gemm(L.ptr(a), L.ptr(w), M, N, K, L.ptr(bias), L.ptr(res), 0, L.ptr(o32), None, 0, L.cur_stream()),
        "fp32+fp16 out + residual": lambda: lib.dfb_gemm(L.ptr(a), L.ptr(w), M, N, K, L.ptr(bias), L.ptr(res), 0, L.ptr(o32), L.ptr(o16), 0, L.cur_stream()),
        "fp32+fp16 out + residual + LN partials": lambda: lib.dfb_gemm_stats(L.ptr(a), L.ptr(w), M, N, K, L.ptr(bias), L.ptr(res), L.ptr(o32), L.ptr(o16), 0, L.ptr(stats), C.byref(tn), L.cur_stream()),
    }
    for name, fn in variants.items():
        L.check(fn()); torch.cuda.synchronize()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(20): fn()
            g.replay(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): g.replay()
            e1.record(); torch.cuda.synchronize()
        print(f"   {name:40s}: {e0.elapsed_time(e1) * 1e3 / 200:6.2f} us per launch")
else:
    for skip, what in [(0, "normal"), (4, "no global stores"), (2, "no tile finish (phase 2)"), (3, "no staging, no finish")]:
        print(f"DFB_DEBUG_SKIP={skip}: {what}", flush=True)
        subprocess.run([sys.executable, "tools/microbench_epilogue_cost.py", "child"] + sys.argv[1:], env=dict(os.environ, DFB_DEBUG_SKIP=str(skip)), timeout=120)
