"""Diagnostic for the tcgen05 GEMM on a real GPU: prints where (rows / columns / k-blocks) a wrong
result deviates, which identifies descriptor / swizzle / TMEM-lane mistakes quickly."""
import math
import sys

import torch

sys.path.insert(0, ".")
from diff_foley_b200 import _lib as L

DEV = "cuda"


def run(M, N, K, splits=1, pattern="rand"):
    g = torch.Generator().manual_seed(0)
    if pattern == "rand":
        a = torch.randn(M, K, generator=g)
        w = torch.randn(N, K, generator=g) / math.sqrt(K)
    else:  # structured: a[m,k] = (m+1) if k==0 ; w[n,k] = (n+1) if k==0
        a = torch.zeros(M, K); w = torch.zeros(N, K)
        a[:, 0] = torch.arange(1, M + 1).float() / 64
        w[:, 0] = torch.arange(1, N + 1).float() / 64
    a = a.to(DEV).half(); w = w.to(DEV).half()
    out = torch.full((M, N), float("nan"), device=DEV)
    rc = L.lib().dfb_gemm(L.ptr(a), L.ptr(w), M, N, K, None, None, 0, L.ptr(out), None, splits, L.cur_stream())
    if rc:
        print("rc", rc, L.lib().dfb_last_error()); return
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t()
    err = (out - ref).abs()
    nan = torch.isnan(out).sum().item()
    rel = float((out - ref).norm() / ref.norm()) if nan == 0 else float("nan")
    print(f"M={M} N={N} K={K} splits={splits} {pattern}: rel={rel:.3e} max_err={err.nan_to_num(1e9).max().item():.3e} nans={nan}")
    if not (rel < 1e-5):
        bad_rows = (err.nan_to_num(1e9) > 1e-3).any(1).nonzero().flatten().tolist()
        bad_cols = (err.nan_to_num(1e9) > 1e-3).any(0).nonzero().flatten().tolist()
        print("  bad rows", bad_rows[:40], "n=", len(bad_rows))
        print("  bad cols", bad_cols[:40], "n=", len(bad_cols))
        print("  out[0,:8]", out[0, :8].tolist())
        print("  ref[0,:8]", ref[0, :8].tolist())
        print("  out[1,:8]", out[1, :8].tolist())
        print("  ref[1,:8]", ref[1, :8].tolist())
        print("  out[9,:8]", out[min(9, M - 1), :8].tolist())
        print("  ref[9,:8]", ref[min(9, M - 1), :8].tolist())


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), L.lib().dfb_version())
    run(128, 128, 64, 1, "struct")
    run(128, 128, 64, 1, "rand")
    run(128, 64, 64, 1, "rand")
    run(128, 128, 128, 1, "rand")
    run(128, 128, 1024, 1, "rand")
    run(256, 256, 512, 1, "rand")
    run(100, 320, 320, 1, "rand")
    run(128, 128, 1024, 4, "rand")
