"""stem (conv3x3 4 -> 320 on the NCHW latent) and head (conv3x3 320 -> 4) kernels in isolation, graph-replayed."""
import sys
sys.path.insert(0, ".")
import torch
from diff_foley_b200 import _lib as L
dev = "cuda"; lib = L.lib()
B, H, W, C = (int(sys.argv[1]) if len(sys.argv) > 1 else 2), 16, 64, 320
x = torch.randn(B, 4, H, W, device=dev); ws = torch.randn(36, C, device=dev) * 0.1; bs = torch.randn(C, device=dev)
o = torch.empty(B * H * W, C, device=dev)
a = torch.randn(B, H, W, C, device=dev).half(); wh = torch.randn(4, 9, C, device=dev) * 0.05; bh = torch.randn(4, device=dev)
oh = torch.empty(B, 4, H, W, device=dev)
fns = {"stem": lambda: L.check(lib.dfb_stem_conv(L.ptr(x), B, 4, H, W, L.ptr(ws), L.ptr(bs), C, L.ptr(o), L.cur_stream())),
       "head": lambda: L.check(lib.dfb_head_conv(L.ptr(a), B, H, W, C, L.ptr(wh), L.ptr(bh), 4, L.ptr(oh), L.cur_stream()))}
for name, fn in fns.items():
    fn(); torch.cuda.synchronize()
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
    print(f"{name} B={B}: {e0.elapsed_time(e1) * 1e3 / 200:6.2f} us per launch")
# reference check (fp32 conv)
import torch.nn.functional as F
ref = F.conv2d(x, ws.t().reshape(C, 4, 3, 3), bs, padding=1).permute(0, 2, 3, 1).reshape(B * H * W, C)
print("stem max err", float((o - ref).abs().max()))
refh = F.conv2d(a.float().permute(0, 3, 1, 2), wh.permute(0, 2, 1).reshape(4, C, 3, 3), bh, padding=1)
print("head max err", float((oh - refh).abs().max()))
