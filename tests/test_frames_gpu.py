"""GPU parity of the CAVP frame ingest (row N4): `dfb_frames_resize` against the oracle restatement of Pillow's
resample (bit-exact: integer arithmetic) and against Pillow itself, and the windowed feature extractor against
encoding its windows one by one."""
import numpy as np
import pytest
import torch

from diff_foley_b200.frames import preprocess_frames
from oracle import frames_oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,H,W", [(3, 360, 640), (2, 100, 150), (1, 224, 224), (2, 480, 227), (5, 37, 53), (32, 270, 480)])
def test_frames_resize_bit_exact(N, H, W):
    rng = np.random.default_rng(N * 7 + H + W)
    frames = rng.integers(0, 256, (N, H, W, 3), dtype=np.uint8)
    out, out8 = preprocess_frames(frames, (224, 224), bgr=True, return_u8=True)
    torch.cuda.synchronize()
    assert out.shape == (N, 3, 224, 224) and out.dtype == torch.float32
    for i in range(min(N, 3)):
        want = frames_oracle.preprocess_frame(frames[i])
        assert np.array_equal(out[i].cpu().numpy(), want), f"frame {i}"
        assert np.array_equal(out8[i].cpu().numpy(), frames_oracle.resize_bilinear_u8(np.ascontiguousarray(frames[i][:, :, ::-1]), 224, 224))
    # size-independent property: a constant frame stays constant (the fixed-point weights of every output sample sum
    # to 2^22 +- rounding; Pillow itself has the same +-1 behaviour only through clipping, which a constant cannot hit)
    const = np.full((1, H, W, 3), 200, dtype=np.uint8)
    o = preprocess_frames(const, (224, 224))
    assert float((o - 200.0 / 255.0).abs().max()) <= 1.0 / 255.0 + 1e-7


def test_frames_resize_matches_pillow():
    from PIL import Image
    import torchvision.transforms as T
    rng = np.random.default_rng(5)
    frames = rng.integers(0, 256, (2, 360, 640, 3), dtype=np.uint8)
    tf = T.Compose([T.Resize((224, 224)), T.ToTensor()])
    want = torch.stack([tf(Image.fromarray(np.ascontiguousarray(f[:, :, ::-1]))) for f in frames])
    got = preprocess_frames(frames, (224, 224), bgr=True).cpu()
    assert torch.equal(got, want)


def test_extract_cavp_features_windows():
    """Extract_CAVP_Features.forward_frames: 11 frames, batch_size 4 -> windows 4 + 4 + 3 (demo_util.py:153-166);
    batching the full windows must give what encoding every window alone gives."""
    from diff_foley_b200.demo_util import Extract_CAVP_Features
    cfg = {"model": {"target": "model.cavp_model.CAVP_Inference",
                     "params": {"video_encode": "Slowonly_pool", "spec_encode": "cnn14_pool", "embed_dim": 512,
                                "video_pretrained": True, "audio_pretrained": True}}}
    ex = Extract_CAVP_Features(fps=4, batch_size=4, device="cuda", config_path=cfg, ckpt_path=None, windows_per_call=2)
    g = torch.Generator().manual_seed(2)
    for _, p in ex.stage1_model.named_parameters():
        if p.dim() > 1:
            p.data.copy_(torch.randn(p.shape, generator=g) * (1.0 / p[0].numel()) ** 0.5)
    rng = np.random.default_rng(3)
    frames = rng.integers(0, 256, (11, 120, 160, 3), dtype=np.uint8)
    feats = ex.forward_frames(frames)
    assert feats.shape == (11, 512) and feats.dtype == np.float32 and np.isfinite(feats).all()
    assert np.allclose(np.linalg.norm(feats, axis=-1), 1.0, atol=2e-3)      # normalize=True
    ex1 = ex
    ex1.windows_per_call = 1
    single = ex1.forward_frames(frames)
    err = np.linalg.norm(feats - single) / np.linalg.norm(single)
    print(f"\n[parity] windows batched vs one by one: rel-L2 {err:.2e}")
    assert err < 2e-3
