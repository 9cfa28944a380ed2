"""Fused attention kernel in isolation: the UNet's self-attention shapes (Lq = Lk, 8 heads) at a given
batch, graph-replayed back-to-back launches, event-timed.  TFLOP/s counts the un-padded 4*Lq*Lk*d.
    python tools/microbench_attention.py [B] [reps]"""
import sys

sys.path.insert(0, ".")
import torch

from diff_foley_b200 import _lib as L

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
dev = "cuda"
lib = L.lib()
heads = 8
print(f"# attention_tcgen05, B={B}, heads={heads}")
for (Lq, Lk, d) in [(1024, 1024, 40), (256, 256, 80), (64, 64, 160), (1024, 32, 40), (256, 32, 80)]:
    dpad = (d + 15) // 16 * 16
    hp = heads * dpad
    qkv = (torch.randn(B * max(Lq, Lk), 3 * hp, device=dev) * 0.5).half()
    out = torch.empty(B * Lq, heads * d, device=dev, dtype=torch.float16)
    scale = d ** -0.5

    def fn():
        L.check(lib.dfb_attention(L.ptr(qkv), 3 * hp, L.ptr(qkv[:, hp:]), 3 * hp, L.ptr(qkv[:, 2 * hp:]), 3 * hp,
                                  L.ptr(out), heads * d, B, heads, Lq, Lk, d, dpad, scale, L.cur_stream()), "att")

    fn()
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(reps):
                fn()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (5 * reps)
    fl = 4.0 * B * heads * Lq * Lk * d
    print(f"Lq={Lq:5d} Lk={Lk:5d} d={d:3d}: {us:8.2f} us  {fl / us / 1e6:8.1f} TFLOP/s")
