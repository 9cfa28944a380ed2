import sys
sys.path.insert(0, ".")
import torch
from diff_foley_b200 import _lib as L
lib = L.lib(); dev = "cuda"
def run(M, N, K, splits, reps=3, act=0):
    g = torch.Generator().manual_seed(1)
    a = torch.randn(M, K, generator=g).to(dev).half(); w = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev).half()
    ref = a.float() @ w.float().t()
    for i in range(reps):
        o = torch.zeros(M, N, device=dev)
        L.check(lib.dfb_gemm(L.ptr(a), L.ptr(w), M, N, K, None, None, act, L.ptr(o), None, splits, L.cur_stream()))
        torch.cuda.synchronize()
        e = (o - ref).abs()
        bad_rows = (e > 1e-3).any(1).nonzero().flatten().tolist()
        bad_cols = (e > 1e-3).any(0).nonzero().flatten().tolist()
        print(f"M={M} N={N} K={K} splits={splits} rep={i} rel={float((o-ref).norm()/ref.norm()):.3e} bad_rows={bad_rows[:8]}..{len(bad_rows)} bad_cols={bad_cols[:8]}..{len(bad_cols)}")
run(128, 128, 1024, 4)
run(32, 1280, 1280, 5)
run(2048, 320, 1280, 3)
run(200, 320, 640, 2)
run(128, 1280, 11520, 14)
