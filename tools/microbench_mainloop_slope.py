"""Steady-state cost of one 128 x BN x 64 k-block per SM: one wave (1 CTA / SM, or 2 with the shallow ring), long K."""
import sys
sys.path.insert(0, ".")
import torch
from diff_foley_b200 import _lib as L
dev = "cuda"; lib = L.lib()
clk = 1.9e9
def run(M, N, K, bn, deep, pair, tag):
    lib.dfb_debug_igemm_force(bn, deep); lib.dfb_debug_igemm_pair(pair)
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
for (tag, M, N, bn, deep, pair) in [("bn128 deep 1cta/sm", 148 * 128, 128, 128, 1, 0), ("bn128 shallow 2cta/sm", 148 * 128, 256, 128, 0, 0),
                                    ("bn128 deep N=256 (2 waves)", 148 * 128, 256, 128, 1, 0),
                                    ("pair128 deep 1cta/sm", 148 * 128, 128, 128, 1, 1), ("pair128 shallow 2cta/sm", 148 * 128, 256, 128, 0, 1),
                                    ("bn64 deep 1cta/sm", 148 * 128, 64, 64, 1, 0), ("bn64 shallow 2cta/sm", 148 * 128, 128, 64, 0, 0)]:
    t1 = run(M, N, 4096, bn, deep, pair, tag); t2 = run(M, N, 12288, bn, deep, pair, tag)
    kb = (12288 - 4096) / 64
    per = (t2 - t1) / kb
    tiles_per_sm = (M / 128) * (N / bn) / 148
    print(f"{tag:28s}: K=4096 {t1:7.1f} us  K=12288 {t2:7.1f} us  -> {per * 1e3:6.1f} ns per k-step = {per * 1e-6 * clk / tiles_per_sm:6.0f} clk per tile-kblock/SM; {2.0 * M * N * 12288 / t2 / 1e6:6.0f} TF/s")
