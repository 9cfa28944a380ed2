"""First-stage decoder on the B200 kernels: latent -> mel-spectrogram image.

`AutoencoderKLDecoderB200` mirrors the decode half of the reference's `AutoencoderKL`
(diff_foley/models/autoencoder.py:285-333: `post_quant_conv` + `Decoder`; Decoder.forward
diff_foley/modules/stage1_autoencoder/model.py:630-663, ResnetBlock :177-242, AttnBlock :245-300,
Upsample :137-152) with the same state-dict keys (`post_quant_conv.*`, `decoder.conv_in.*`,
`decoder.mid.block_1.*`, `decoder.mid.attn_1.{norm,q,k,v,proj_out}.*`, `decoder.up.L.block.I.*`,
`decoder.up.L.upsample.conv.*`, `decoder.norm_out.*`, `decoder.conv_out.*`), so
`load_state_dict(first_stage_sd, strict=False)` of the reference checkpoint fills it, and
`decode(z)` / `decode_first_stage(z)` keep the reference contracts (ddpm.py:739-797:
`z / scale_factor` first).  Config: Stage2_LDM.yaml:38-57.

Compute path -- every convolution / linear on `igemm_tcgen05_kernel`, fp16 channels-last operands,
fp32 accumulation and an fp32 residual stream (the same datapath as the UNet):
  * GroupNorm(32, eps 1e-6) + swish -> fp16 operand (`dfb_groupnorm`; slabs of the upper levels go
    through its statistics + apply kernels);
  * 3x3 convs as 9-tap implicit GEMMs with the residual add in the epilogue (`dfb_conv3x3`); the 1x1
    `nin_shortcut` of the channel-changing blocks is fused into conv2 as extra K columns
    (`dfb_conv3x3_cat`);
  * `post_quant_conv` (1x1, 4->4) is folded EXACTLY into `conv_in`: the latent gets a constant-one
    channel, so the 1x1 bias passes through the 3x3 taps only where the image exists (TMA zero fill
    pads the ones channel like every other channel);
  * the single-head 512-wide attention at 16x64: q / k projections and v^T = W_v h^T as GEMMs, scores
    q k^T as a GEMM, `dfb_softmax_rows`, P v as a GEMM with b_v added after it (softmax rows sum to one),
    proj_out GEMM with the residual in its epilogue;
  * nearest-2x upsample feeds the following conv's fp16 operand directly (`dfb_upsample2x_f16`).
Only the NCHW <-> channels-last conversion at the two ends uses torch ops.  No CPU / PyTorch fallback.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib as L

VAE_CFG = dict(ch=128, out_ch=3, ch_mult=(1, 2, 4, 4), num_res_blocks=2, z_channels=4, embed_dim=4)
SCALE_FACTOR = 0.18215


class _Conv(nn.Module):
    def __init__(self, cin, cout, k):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin, k, k))
        self.bias = nn.Parameter(torch.empty(cout))
        b = 1.0 / math.sqrt(cin * k * k)
        nn.init.uniform_(self.weight, -b, b)
        nn.init.uniform_(self.bias, -b, b)


class _Norm(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))


class _Res(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.norm1, self.conv1 = _Norm(cin), _Conv(cin, cout, 3)
        self.norm2, self.conv2 = _Norm(cout), _Conv(cout, cout, 3)
        if cin != cout:
            self.nin_shortcut = _Conv(cin, cout, 1)


class _Attn(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.norm = _Norm(c)
        self.q, self.k, self.v, self.proj_out = (_Conv(c, c, 1) for _ in range(4))


class _Up(nn.Module):
    pass


class _Decoder(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        ch, mults = cfg["ch"], cfg["ch_mult"]
        block_in = ch * mults[-1]
        self.conv_in = _Conv(cfg["z_channels"], block_in, 3)
        self.mid = nn.Module()
        self.mid.block_1 = _Res(block_in, block_in)
        self.mid.attn_1 = _Attn(block_in)
        self.mid.block_2 = _Res(block_in, block_in)
        ups = [None] * len(mults)
        for lvl in reversed(range(len(mults))):
            up = _Up()
            blocks = []
            for _ in range(cfg["num_res_blocks"] + 1):
                blocks.append(_Res(block_in, ch * mults[lvl]))
                block_in = ch * mults[lvl]
            up.block = nn.ModuleList(blocks)
            if lvl != 0:
                up.upsample = nn.Module()
                up.upsample.conv = _Conv(block_in, block_in, 3)
            ups[lvl] = up
        self.up = nn.ModuleList(ups)
        self.norm_out = _Norm(block_in)
        self.conv_out = _Conv(block_in, cfg["out_ch"], 3)


def _pack3(w):
    """OIHW -> fp16 [N, tap*C + c] (tap = ky*3 + kx)."""
    n, c = w.shape[:2]
    return w.detach().float().permute(0, 2, 3, 1).reshape(n, 9 * c).half().contiguous()


class AutoencoderKLDecoderB200(nn.Module):
    def __init__(self, ddconfig=None, embed_dim=4, scale_factor=SCALE_FACTOR, **ignored):
        super().__init__()
        cfg = dict(VAE_CFG)
        if ddconfig:
            cfg.update(ch=ddconfig["ch"], out_ch=ddconfig["out_ch"], ch_mult=tuple(ddconfig["ch_mult"]),
                       num_res_blocks=ddconfig["num_res_blocks"], z_channels=ddconfig["z_channels"])
            if ddconfig.get("attn_resolutions"):
                raise NotImplementedError("attention at the up levels is not used by Diff-Foley's first stage")
        cfg["embed_dim"] = embed_dim
        if cfg["ch"] % 64 or cfg["z_channels"] + 1 > 64 or cfg["out_ch"] > 16:
            raise NotImplementedError("decoder widths must be multiples of 64 (32 groups of an even width)")
        self.cfg, self.scale_factor = cfg, scale_factor
        self.post_quant_conv = _Conv(embed_dim, cfg["z_channels"], 1)
        self.decoder = _Decoder(cfg)
        self._packed = None

    # ------------------------------------------------------------------ weights -> kernel layouts
    def _apply(self, fn, *a, **k):
        self._packed = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._packed = None
        return super().load_state_dict(*a, **k)

    def _pack(self, dev):
        if self._packed is not None:
            return self._packed
        P = {}
        f32 = lambda t: t.detach().float().to(dev).contiguous()

        def conv3(name, m):
            P[name + ".w"], P[name + ".b"] = _pack3(m.weight).to(dev), f32(m.bias)

        def norm(name, m):
            P[name + ".g"], P[name + ".b"] = f32(m.weight), f32(m.bias)

        def res(name, m):
            norm(name + ".norm1", m.norm1)
            conv3(name + ".conv1", m.conv1)
            norm(name + ".norm2", m.norm2)
            if hasattr(m, "nin_shortcut"):  # conv2 | nin_shortcut as one [N, 9*cout + cin] matrix
                w2 = _pack3(m.conv2.weight)
                ws = m.nin_shortcut.weight.detach().float().flatten(1).half()
                P[name + ".conv2.w"] = torch.cat([w2, ws], dim=1).contiguous().to(dev)
                P[name + ".conv2.b"], P[name + ".nin.b"] = f32(m.conv2.bias), f32(m.nin_shortcut.bias)
            else:
                conv3(name + ".conv2", m.conv2)

        d = self.decoder
        # conv_in o post_quant_conv on [z (E ch) | 1 | zeros] padded to 64 channels
        wc = d.conv_in.weight.detach().double()                   # [N, Z, 3, 3]
        wq = self.post_quant_conv.weight.detach().double()[:, :, 0, 0]  # [Z, E]
        bq = self.post_quant_conv.bias.detach().double()
        E = wq.shape[1]
        w_in = torch.zeros(wc.shape[0], 64, 3, 3, dtype=torch.float64, device=wc.device)
        w_in[:, :E] = torch.einsum("nzyx,ze->neyx", wc, wq)
        w_in[:, E] = torch.einsum("nzyx,z->nyx", wc, bq)
        P["conv_in.w"], P["conv_in.b"] = _pack3(w_in.float()).to(dev), f32(d.conv_in.bias)
        res("mid.block_1", d.mid.block_1)
        a = d.mid.attn_1
        norm("attn.norm", a.norm)
        for n in ("q", "k", "v", "proj_out"):
            m = getattr(a, n)
            P[f"attn.{n}.w"] = m.weight.detach().float().flatten(1).half().contiguous().to(dev)
            P[f"attn.{n}.b"] = f32(m.bias)
        res("mid.block_2", d.mid.block_2)
        for lvl, up in enumerate(d.up):
            for i, blk in enumerate(up.block):
                res(f"up.{lvl}.block.{i}", blk)
            if hasattr(up, "upsample"):
                conv3(f"up.{lvl}.upsample", up.upsample.conv)
        norm("norm_out", d.norm_out)
        wo = torch.zeros(16, *d.conv_out.weight.shape[1:], device=d.conv_out.weight.device)  # N padded to 16
        wo[: self.cfg["out_ch"]] = d.conv_out.weight.detach().float()
        bo = torch.zeros(16, device=d.conv_out.bias.device)
        bo[: self.cfg["out_ch"]] = d.conv_out.bias.detach().float()
        P["conv_out.w"], P["conv_out.b"] = _pack3(wo).to(dev), bo.to(dev)
        self._packed = P
        return P

    # -------------------------------------------------------------------------------- kernels
    @staticmethod
    def _gn(x, g, b, silu, raw=False):
        """fp32 [B,H,W,C] -> fp16 GroupNorm(32, eps 1e-6) (+swish) (+ the raw fp16 copy)."""
        B, H, W, C = x.shape
        out = torch.empty(B, H, W, C, device=x.device, dtype=torch.float16)
        rw = torch.empty_like(out) if raw else None
        L.check(L.lib().dfb_groupnorm(L.ptr(x), C, None, 0, B, H * W, L.ptr(g), L.ptr(b), 1e-6, int(silu),
                                      L.ptr(out), L.ptr(rw), L.cur_stream()), "dfb_groupnorm")
        return (out, rw) if raw else out

    @staticmethod
    def _conv(a16, w, bias, N, residual=None):
        B, H, W, C = a16.shape
        out = torch.empty(B, H, W, N, device=a16.device, dtype=torch.float32)
        L.check(L.lib().dfb_conv3x3(L.ptr(a16), L.ptr(w), B, H, W, C, N, L.ptr(bias), None, L.ptr(residual), 0,
                                    L.ptr(out), None, 0, L.cur_stream()), "dfb_conv3x3")
        return out

    @staticmethod
    def _gemm(a16, w, bias=None, residual=None, f16=False):
        M, K = a16.shape
        N = w.shape[0]
        out = torch.empty(M, N, device=a16.device, dtype=torch.float16 if f16 else torch.float32)
        L.check(L.lib().dfb_gemm(L.ptr(a16), L.ptr(w), M, N, K, L.ptr(bias), L.ptr(residual), 0,
                                 None if f16 else L.ptr(out), L.ptr(out) if f16 else None, 0, L.cur_stream()),
                "dfb_gemm")
        return out

    def _res(self, P, name, x):
        cout = P[name + ".conv1.b"].numel()
        fused = (name + ".nin.b") in P
        if fused:
            a, raw = self._gn(x, P[name + ".norm1.g"], P[name + ".norm1.b"], True, raw=True)
        else:
            a = self._gn(x, P[name + ".norm1.g"], P[name + ".norm1.b"], True)
        h = self._conv(a, P[name + ".conv1.w"], P[name + ".conv1.b"], cout)
        a = self._gn(h, P[name + ".norm2.g"], P[name + ".norm2.b"], True)
        if not fused:
            return self._conv(a, P[name + ".conv2.w"], P[name + ".conv2.b"], cout, residual=x)
        B, H, W, C = a.shape
        out = torch.empty(B, H, W, cout, device=a.device, dtype=torch.float32)
        L.check(L.lib().dfb_conv3x3_cat(L.ptr(a), L.ptr(raw), raw.shape[-1], L.ptr(P[name + ".conv2.w"]), B, H, W, C,
                                        cout, L.ptr(P[name + ".conv2.b"]), L.ptr(P[name + ".nin.b"]), None,
                                        L.ptr(out), None, 0, L.cur_stream()), "dfb_conv3x3_cat")
        return out

    def _attn(self, P, x):
        B, H, W, C = x.shape
        Lq = H * W
        hn = self._gn(x, P["attn.norm.g"], P["attn.norm.b"], False).view(B * Lq, C)
        q = self._gemm(hn, P["attn.q.w"], P["attn.q.b"], f16=True)
        k = self._gemm(hn, P["attn.k.w"], P["attn.k.b"], f16=True)
        o = torch.empty(B * Lq, C, device=x.device, dtype=torch.float16)
        for b in range(B):
            rows = slice(b * Lq, (b + 1) * Lq)
            vT = self._gemm(P["attn.v.w"], hn[rows], f16=True)                    # [C, Lq] = W_v h^T
            s = self._gemm(q[rows], k[rows])                                       # [Lq, Lq] fp32 scores
            p = torch.empty(Lq, Lq, device=x.device, dtype=torch.float16)
            L.check(L.lib().dfb_softmax_rows(L.ptr(s), Lq, Lq, float(C) ** -0.5, L.ptr(p), L.cur_stream()),
                    "dfb_softmax_rows")
            o[rows] = self._gemm(p, vT, P["attn.v.b"], f16=True)                   # P v + b_v
        return self._gemm(o, P["attn.proj_out.w"], P["attn.proj_out.b"], residual=x.view(B * Lq, C)).view(B, H, W, C)

    # ---------------------------------------------------------------------------------- API
    @torch.no_grad()
    def decode(self, z):
        """`AutoencoderKL.decode` (autoencoder.py:330-333): z [B,E,h,w] (already divided by the scale
        factor) -> image [B,out_ch,8h,8w] fp32.  The ~110 launches of one decode are captured once per latent
        shape as a CUDA graph over a static input buffer and replayed (DFB_NO_VAE_GRAPH=1: eager)."""
        if not z.is_cuda:
            raise RuntimeError("AutoencoderKLDecoderB200 runs on a B200 only: there is no CPU path")
        import os
        if os.environ.get("DFB_NO_VAE_GRAPH"):
            return self._decode_eager(z)
        dev = z.device
        packed = self._pack(dev)
        key = (tuple(z.shape), str(dev), id(packed))
        ent = self.__dict__.setdefault("_graphs", {}).get(key)
        if ent is None:
            with torch.cuda.device(dev):
                sz = z.detach().float().clone()
                self._decode_eager(sz)                    # warm-up: lazy workspaces / kernel attributes
                torch.cuda.synchronize(dev)
                graph = torch.cuda.CUDAGraph()
                side = torch.cuda.Stream(dev)
                side.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(side):
                    with torch.cuda.graph(graph, stream=side):
                        out = self._decode_eager(sz)
                torch.cuda.current_stream(dev).wait_stream(side)
            ent = self._graphs[key] = (graph, sz, out)
        graph, sz, out = ent
        sz.copy_(z)
        graph.replay()
        return out.clone()

    @torch.no_grad()
    def _decode_eager(self, z):
        P = self._pack(z.device)
        B, E, H, W = z.shape
        with torch.cuda.device(z.device):
            zin = torch.zeros(B, H, W, 64, device=z.device, dtype=torch.float16)
            zin[..., :E] = z.permute(0, 2, 3, 1)
            zin[..., E] = 1.0
            h = self._conv(zin, P["conv_in.w"], P["conv_in.b"], P["conv_in.b"].numel())
            h = self._res(P, "mid.block_1", h)
            h = self._attn(P, h)
            h = self._res(P, "mid.block_2", h)
            for lvl in reversed(range(len(self.cfg["ch_mult"]))):
                for i in range(self.cfg["num_res_blocks"] + 1):
                    h = self._res(P, f"up.{lvl}.block.{i}", h)
                if lvl != 0:
                    B_, H_, W_, C_ = h.shape
                    a = torch.empty(B_, 2 * H_, 2 * W_, C_, device=h.device, dtype=torch.float16)
                    L.check(L.lib().dfb_upsample2x_f16(L.ptr(h), L.ptr(a), B_, H_, W_, C_, L.cur_stream()),
                            "dfb_upsample2x_f16")
                    h = self._conv(a, P[f"up.{lvl}.upsample.w"], P[f"up.{lvl}.upsample.b"], C_)
            a = self._gn(h, P["norm_out.g"], P["norm_out.b"], True)
            img = self._conv(a, P["conv_out.w"], P["conv_out.b"], 16)
        return img[..., : self.cfg["out_ch"]].permute(0, 3, 1, 2).contiguous()

    def decode_first_stage(self, z):
        """`LatentDiffusion.decode_first_stage` (ddpm.py:739-797): `1/scale_factor * z`, then decode."""
        return self.decode(z / self.scale_factor)

    forward = decode
