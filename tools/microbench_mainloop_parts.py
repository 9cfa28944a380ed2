"""Which side bounds the main loop of a 1-CTA/SM (deep ring) tile: per-k-block time with parts disabled.

The DFB_DEBUG_SKIP bits 32 (no MMA instructions), 64 (no activation loads), 128 (no weight loads) this script toggles
were a temporary instrumentation of the v8 kernel (they cost ~10 % in the hot loops and were removed with the
rewrite of those loops); profiles/r2_mainloop_bound_before_elect.log is its output on that build.  On the current
kernel only the un-instrumented line (`python tools/microbench_mainloop_parts.py child`) is meaningful."""
import os, subprocess, sys
if len(sys.argv) > 1:
    sys.path.insert(0, ".")
    import torch
    from diff_foley_b200 import _lib as L
    dev = "cuda"; lib = L.lib()
    def run(M, N, K, bn, deep):
        lib.dfb_debug_igemm_force(bn, deep)
        a = torch.randn(M, K, device=dev).half(); w = (torch.randn(N, K, device=dev) / K ** 0.5).half()
        out = torch.empty(M, N, device=dev, dtype=torch.float16)
        fn = lambda: L.check(lib.dfb_gemm(L.ptr(a), L.ptr(w), M, N, K, None, None, 0, None, L.ptr(out), 1, L.cur_stream()), "gemm")
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / 10
    for (tag, M, N, bn, deep) in [("bn128 deep 1cta/sm", 148 * 128, 128, 128, 1), ("bn128 shallow 2cta/sm", 148 * 128, 256, 128, 0), ("bn64 deep 1cta/sm", 148 * 128, 64, 64, 1),
                                  ("M=32 bn128 (a32 if on) 80 ctas", 32, 1280 , 128, 1)]:
        if M == 32:
            # split-K forced through the table path is not available here: emulate one CTA's stream with 10 n-tiles, long K
            t1 = run(M, N, 8192, bn, deep); t2 = run(M, N, 24576, bn, deep); kb = (24576 - 8192) / 64; tiles = 1
        else:
            t1 = run(M, N, 4096, bn, deep); t2 = run(M, N, 12288, bn, deep); kb = (12288 - 4096) / 64
            tiles = (M / 128) * (N / bn) / 148
        per = (t2 - t1) / kb
        print(f"   {tag:32s}: {per * 1e3:6.1f} ns per k-step -> {per * 1e3 / tiles:6.1f} ns per tile-kblock per SM")
else:
    for skip, what in [(0, "normal"), (32, "no MMA instructions (TMA pipeline alone)"), 
                       (64 + 32, "no MMA, no A loads"), (128 + 32, "no MMA, no W loads"), (64 + 128 + 32, "no MMA, no loads (barrier round trips only)")]:
        print(f"DFB_DEBUG_SKIP={skip}: {what}", flush=True)
        env = dict(os.environ, DFB_DEBUG_SKIP=str(skip))
        subprocess.run([sys.executable, "tools/microbench_mainloop_parts.py", "child"], env=env, timeout=120)
