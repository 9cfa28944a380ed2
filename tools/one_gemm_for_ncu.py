"""One B_eff = 16 level-0 conv (or a Linear) in isolation, for ncu: python tools/one_gemm_for_ncu.py [conv|geglu|lin] [reps]."""
import sys
sys.path.insert(0, ".")
import torch
from diff_foley_b200 import _lib as L
kind = sys.argv[1] if len(sys.argv) > 1 else "conv"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = "cuda"; lib = L.lib()
if kind == "conv":
    B, H, W, C, N = 16, 16, 64, 320, 320
    a = torch.randn(B, H, W, C, device=dev).half()
    w = (torch.randn(N, 9 * C, device=dev) / (9 * C) ** 0.5).half()
    bias = torch.randn(N, device=dev)
    out = torch.empty(B * H * W, N, device=dev)
    fn = lambda: L.check(lib.dfb_conv3x3(L.ptr(a), L.ptr(w), B, H, W, C, N, L.ptr(bias), None, None, 0, L.ptr(out), None, 0, L.cur_stream()), "conv")
    fl = 2.0 * B * H * W * N * 9 * C
else:
    M, N, K = (16384, 1152, 320) if kind == "lin" else (16384, 320, 1600)
    a = torch.randn(M, K, device=dev).half()
    w = (torch.randn(N, K, device=dev) / K ** 0.5).half()
    bias = torch.randn(N, device=dev)
    out = torch.empty(M, N, device=dev, dtype=torch.float16)
    fn = lambda: L.check(lib.dfb_gemm(L.ptr(a), L.ptr(w), M, N, K, L.ptr(bias), None, 0, None, L.ptr(out), 0, L.cur_stream()), "gemm")
    fl = 2.0 * M * N * K
for _ in range(reps): fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): fn()
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / 20
print(f"{kind}: {us:.2f} us  {fl / us / 1e6:.1f} TFLOP/s")
