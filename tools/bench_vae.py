"""First-stage decode timing (row a19 / N1): ms per clip and TFLOP/s, B = 1 and 8 clips.
    python tools/bench_vae.py"""
import sys

sys.path.insert(0, ".")
import torch

from diff_foley_b200.vae import AutoencoderKLDecoderB200

GFLOP_PER_CLIP = 620.0
vae = AutoencoderKLDecoderB200().cuda()
for B in (1, 8):
    z = torch.randn(B, 4, 16, 64, device="cuda")
    for _ in range(3):
        vae.decode_first_stage(z)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
        vae.decode_first_stage(z)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"first-stage decode B={B}: {ms:7.2f} ms ({ms / B:6.2f} ms/clip, {B * GFLOP_PER_CLIP / ms:6.0f} TFLOP/s... x1e-3)"
          .replace("TFLOP/s... x1e-3", "GFLOP/ms = TFLOP/s"))
