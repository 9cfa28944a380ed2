"""Why do kernels run ~2x slower inside the UNet plan than back-to-back in isolation?  Graph-replayed
sequences that vary one factor at a time (descriptor reuse, cold weights, kernel interleaving)."""
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

def gemm_fn(a, w, o, res=None, splits=1):
    M, K = a.shape; N = w.shape[0]
    def f():
        L.check(lib.dfb_gemm(L.ptr(a), L.ptr(w), M, N, K, None, L.ptr(res), 0, L.ptr(o), None, splits, L.cur_stream()))
    return f

def ln_fn(x, o):
    rows, C = x.shape
    gm = torch.ones(C, device=dev); bt = torch.zeros(C, device=dev)
    def f(): L.check(lib.dfb_layernorm(L.ptr(x), rows, C, L.ptr(gm), L.ptr(bt), 1e-5, L.ptr(o), L.cur_stream()))
    return f

def gn_fn(x, o, B, HW, C):
    gm = torch.ones(C, device=dev); bt = torch.zeros(C, device=dev)
    def f(): L.check(lib.dfb_groupnorm(L.ptr(x), C, None, 0, B, HW, L.ptr(gm), L.ptr(bt), 1e-5, 1, L.ptr(o), None, L.cur_stream()))
    return f

M, N, K = 2048, 320, 320
n = 64
As = [torch.randn(M, K, device=dev).half() for _ in range(n)]
Ws = [(torch.randn(N, K, device=dev) / K ** 0.5).half() for _ in range(n)]
Os = [torch.empty(M, N, device=dev) for _ in range(n)]
print("same buffers           :", f"{run_graph([gemm_fn(As[0], Ws[0], Os[0])] * n):.2f} us")
print("64 distinct A/W/out    :", f"{run_graph([gemm_fn(As[i], Ws[i], Os[i]) for i in range(n)]):.2f} us")
print("distinct W only        :", f"{run_graph([gemm_fn(As[0], Ws[i], Os[0]) for i in range(n)]):.2f} us")
print("distinct A only        :", f"{run_graph([gemm_fn(As[i], Ws[0], Os[0]) for i in range(n)]):.2f} us")
print("distinct out only      :", f"{run_graph([gemm_fn(As[0], Ws[0], Os[i]) for i in range(n)]):.2f} us")
# chain: out of LN feeds the GEMM (producer->consumer through L2), like the real plan
X = torch.randn(M, K, device=dev); A16 = torch.empty(M, K, device=dev, dtype=torch.float16)
seq = []
for i in range(n):
    seq += [ln_fn(X, A16), gemm_fn(A16, Ws[i], Os[i % 4])]
t_pair = run_graph(seq) * 2
print("LN -> GEMM pairs       :", f"{t_pair:.2f} us per pair")
print("LN alone               :", f"{run_graph([ln_fn(X, A16)] * n):.2f} us")
# big cold weights: 29.5 MB each, 24 distinct (708 MB > L2)
M2, N2, K2 = 128, 1280, 11520
A2 = torch.randn(M2, K2, device=dev).half()
W2 = [(torch.randn(N2, K2, device=dev) / K2 ** 0.5).half() for _ in range(24)]
O2 = torch.empty(M2, N2, device=dev)
print("128x1280x11520 s14 warm:", f"{run_graph([gemm_fn(A2, W2[0], O2, splits=14)] * 24):.2f} us")
print("128x1280x11520 s14 cold:", f"{run_graph([gemm_fn(A2, W2[i], O2, splits=14) for i in range(24)]):.2f} us   (29.5 MB each)")
M3, N3, K3 = 128, 1280, 1280
A3 = torch.randn(M3, K3, device=dev).half()
W3 = [(torch.randn(N3, K3, device=dev) / K3 ** 0.5).half() for _ in range(64)]
O3 = torch.empty(M3, N3, device=dev)
print("128x1280x1280 s5 warm  :", f"{run_graph([gemm_fn(A3, W3[0], O3, splits=5)] * 64):.2f} us")
print("128x1280x1280 s5 cold  :", f"{run_graph([gemm_fn(A3, W3[i], O3, splits=5) for i in range(64)]):.2f} us   (3.3 MB each, 210 MB pool)")
print("128x1280x1280 s1 cold  :", f"{run_graph([gemm_fn(A3, W3[i], O3, splits=1) for i in range(64)]):.2f} us")
print("128x1280x1280 s2 cold  :", f"{run_graph([gemm_fn(A3, W3[i], O3, splits=2) for i in range(64)]):.2f} us")
# GN in isolation vs interleaved
Xg = torch.randn(2, 1024, 320, device=dev); Og = torch.empty(2, 1024, 320, device=dev, dtype=torch.float16)
print("GN 2x1024x320 alone    :", f"{run_graph([gn_fn(Xg, Og, 2, 1024, 320)] * n):.2f} us")
