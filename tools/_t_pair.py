import sys, math
sys.path.insert(0, ".")
import torch
from diff_foley_b200 import _lib as L
lib = L.lib()
dev = "cuda"
for (M, N, K, res) in [(512, 256, 128, 0), (2048, 320, 320, 1), (300, 128, 64, 0), (4096, 640, 640, 1), (16384, 320, 1600, 1)]:
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).to(dev).half()
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(dev).half()
    b = torch.randn(N, generator=g).to(dev)
    r = torch.randn(M, N, generator=g).to(dev) if res else None
    out = torch.full((M, N), float("nan"), device=dev)
    L.check(lib.dfb_gemm(L.ptr(a), L.ptr(w), M, N, K, L.ptr(b), L.ptr(r), 0, L.ptr(out), None, 1, L.cur_stream()), "gemm")
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t() + b + (r if res else 0)
    err = float((out - ref).norm() / ref.norm())
    print(f"M={M} N={N} K={K}: rel-L2 {err:.2e} finite={bool(torch.isfinite(out).all())}", flush=True)
