"""Oracle (TEST INFRASTRUCTURE ONLY): CPU restatement of the frame preprocessing of
`Extract_CAVP_Features.forward` (reference inference/demo_util.py:147-152):

    rgb = cv2.cvtColor(rgb, cv2.COLOR_BGR2RGB)
    rgb_tensor = transforms.Compose([Resize((224, 224)), ToTensor()])(Image.fromarray(rgb))

The arithmetic lives in a third-party dependency that is not vendored in the reference tree: Pillow
(requirements.txt pins `Pillow==9.4.0`; this image has 12.2.0 -- the 8-bit resample path below is unchanged
since Pillow 3.0's ImagingResample rewrite).  Published algorithm restated from Pillow's
src/libImaging/Resample.c: bilinear_filter, precompute_coeffs, normalize_coeffs_8bpc,
ImagingResampleHorizontal_8bpc, ImagingResampleVertical_8bpc.  Pinned by tests/test_oracle.py against Pillow
itself (Image.resize(..., BILINEAR) + torchvision ToTensor) on up-, down- and same-size cases, bit for bit.
Plain numpy integer arithmetic, loops over the output axis only.
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def precompute_coeffs(in_size, out_size):
    """Resample.c precompute_coeffs (box = the whole axis) + normalize_coeffs_8bpc for the bilinear (triangle)
    filter, support 1.0.  Returns (kk int32 [out, ksize], bounds int32 [out, 2] = (xmin, count), ksize)."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = np.zeros(xmax, dtype=np.float64)
        for x in range(xmax):
            a = abs((x + xmin - center + 0.5) * ss)
            w[x] = 1.0 - a if a < 1.0 else 0.0
        ww = w.sum()
        if ww != 0.0:
            w = w / ww
        for x in range(xmax):
            kk[xx, x] = int(-0.5 + w[x] * (1 << PRECISION_BITS)) if w[x] < 0 else int(0.5 + w[x] * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return kk, bounds, ksize


def _clip8(v):
    return np.clip(v >> PRECISION_BITS, 0, 255).astype(np.uint8)


def resize_bilinear_u8(img, out_h, out_w):
    """img uint8 [H, W, 3] -> uint8 [out_h, out_w, 3]: horizontal pass, round to uint8, vertical pass."""
    H, W, _ = img.shape
    kh, bh, _ = precompute_coeffs(W, out_w)
    tmp = np.empty((H, out_w, 3), dtype=np.uint8)
    src = img.astype(np.int64)
    for ox in range(out_w):
        x0, n = bh[ox]
        acc = (1 << (PRECISION_BITS - 1)) + (src[:, x0:x0 + n, :] * kh[ox, :n].astype(np.int64)[None, :, None]).sum(1)
        tmp[:, ox, :] = _clip8(acc)
    kv, bv, _ = precompute_coeffs(H, out_h)
    out = np.empty((out_h, out_w, 3), dtype=np.uint8)
    src = tmp.astype(np.int64)
    for oy in range(out_h):
        y0, n = bv[oy]
        acc = (1 << (PRECISION_BITS - 1)) + (src[y0:y0 + n] * kv[oy, :n].astype(np.int64)[:, None, None]).sum(0)
        out[oy] = _clip8(acc)
    return out


def preprocess_frame(bgr, out_h=224, out_w=224):
    """demo_util.py:148-149 for one cv2 frame: uint8 BGR [H,W,3] -> float32 [3,out_h,out_w] in [0,1]."""
    rgb = bgr[:, :, ::-1]
    r = resize_bilinear_u8(np.ascontiguousarray(rgb), out_h, out_w)
    return (r.astype(np.float32) / np.float32(255.0)).transpose(2, 0, 1)
