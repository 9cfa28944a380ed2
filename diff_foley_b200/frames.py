"""GPU frame ingest for the CAVP video encoder (reference inference/demo_util.py:135-163).

The reference converts every decoded frame on the CPU -- cv2.cvtColor(BGR2RGB), PIL Resize((224, 224)),
ToTensor -- and copies one fp32 [1,3,224,224] tensor to the GPU per frame.  Here a whole window of uint8
frames is copied once (3 bytes per source pixel) and `dfb_frames_resize` does channel swap + Pillow's
two-pass 8-bit antialiased bilinear resample + the /255 of ToTensor in two launches, bit-identical to the
reference's preprocessing for every source size.  This module holds the host side: Pillow's coefficient
tables (Resample.c: precompute_coeffs for the triangle filter + normalize_coeffs_8bpc), cached per axis.
"""
import math

import numpy as np
import torch

from . import _lib as L

_PRECISION_BITS = 32 - 8 - 2
_tables = {}


def _bilinear_tables(in_size, out_size):
    """int32 fixed-point coefficients [out, ksize] and (first index, count) bounds [out, 2] of Pillow's
    bilinear resample from `in_size` to `out_size` samples (whole-axis box)."""
    scale = float(in_size) / out_size
    fscale = max(scale, 1.0)
    support = fscale                                     # bilinear support = 1.0
    ksize = int(math.ceil(support)) * 2 + 1
    xx = np.arange(out_size, dtype=np.float64)
    center = (xx + 0.5) * scale
    xmin = np.maximum((center - support + 0.5).astype(np.int64), 0)      # C int cast: values are >= -0.5 -> 0
    xmax = np.minimum((center + support + 0.5).astype(np.int64), in_size) - xmin
    x = np.arange(ksize, dtype=np.float64)[None, :]
    w = 1.0 - np.abs((x + xmin[:, None] - center[:, None] + 0.5) / fscale)
    w = np.where((w > 0.0) & (x < xmax[:, None]), w, 0.0)
    ww = w.sum(1, keepdims=True)
    w = np.where(ww != 0.0, w / np.where(ww == 0.0, 1.0, ww), w)
    kk = (0.5 + w * (1 << _PRECISION_BITS)).astype(np.int64).astype(np.int32)   # all weights >= 0 for this filter
    bounds = np.stack([xmin, xmax], 1).astype(np.int32)
    return np.ascontiguousarray(kk), np.ascontiguousarray(bounds), ksize


def resize_tables(in_size, out_size, device):
    key = (in_size, out_size, str(device))
    if key not in _tables:
        kk, bounds, ksize = _bilinear_tables(in_size, out_size)
        _tables[key] = (torch.from_numpy(kk).to(device), torch.from_numpy(bounds).to(device), ksize)
    return _tables[key]


@torch.no_grad()
def preprocess_frames(frames_u8, out_hw=(224, 224), bgr=True, device=None, return_u8=False):
    """frames_u8: uint8 [N,H,W,3] (numpy or torch, host or device; cv2's BGR order when bgr=True)
    -> float32 [N,3,224,224] in [0,1] on the GPU == stack(img_transform(Image.fromarray(rgb)))."""
    t = torch.as_tensor(np.ascontiguousarray(frames_u8) if isinstance(frames_u8, np.ndarray) else frames_u8)
    if t.dtype != torch.uint8 or t.dim() != 4 or t.shape[-1] != 3:
        raise ValueError(f"frames must be uint8 [N,H,W,3], got {t.dtype} {tuple(t.shape)}")
    dev = torch.device(device) if device is not None else (t.device if t.is_cuda else torch.device("cuda"))
    if dev.type != "cuda":
        raise RuntimeError("preprocess_frames runs on a CUDA (sm_100a) device only")
    if not t.is_cuda:
        t = t.pin_memory().to(dev, non_blocking=True)
    t = t.contiguous()
    N, H, W, _ = t.shape
    OH, OW = out_hw
    kh, bh, ksh = resize_tables(W, OW, dev)
    kv, bv, ksv = resize_tables(H, OH, dev)
    tmp = torch.empty(N, H, OW, 3, dtype=torch.uint8, device=dev)
    out = torch.empty(N, 3, OH, OW, dtype=torch.float32, device=dev)
    out8 = torch.empty(N, OH, OW, 3, dtype=torch.uint8, device=dev) if return_u8 else None
    with torch.cuda.device(dev):
        L.check(L.lib().dfb_frames_resize(L.ptr(t), N, H, W, 1 if bgr else 0, L.ptr(kh), L.ptr(bh), ksh, OW, L.ptr(kv),
                                          L.ptr(bv), ksv, OH, L.ptr(tmp), L.ptr(out), L.ptr(out8), L.cur_stream()),
                "dfb_frames_resize")
    return (out, out8) if return_u8 else out
