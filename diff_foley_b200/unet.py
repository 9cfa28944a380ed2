"""UNetModelB200 -- drop-in for the reference's UNetModel
(diff_foley/modules/diffusionmodules/openai_unetmodel.py:413-742).

Same constructor keywords, same parameter names and shapes (so `load_state_dict` of
ldm_epoch240.ckpt's `model.diffusion_model.*` keys works unchanged, demo_util.py:182-184), same
`forward(x, timesteps, context)` contract; select it by pointing `unet_config.target` of
inference/config/Stage2_LDM.yaml at `diff_foley_b200.unet.UNetModelB200`.

The nn.Module tree below only *holds* parameters.  All computation happens in libdfb.so
(include/dfb.h): on first use the parameters are handed to the engine, which packs them into fp16
tensor-core layouts; forward() is one C call that enqueues the launch plan on the current stream.
There is no PyTorch compute path and no fallback: without a B200 the call raises.
"""
import ctypes as C
import math

import torch
import torch.nn as nn

from . import _lib as L


class _Holder(nn.Module):
    """weight (+ bias) of a conv / linear / norm layer, named like the torch.nn original."""

    def __init__(self, wshape, bias=True, kind="linear"):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(*wshape))
        self.bias = nn.Parameter(torch.empty(wshape[0])) if bias else None
        if kind == "norm":
            nn.init.ones_(self.weight)
            nn.init.zeros_(self.bias)
        elif kind == "zero":  # zero_module(...) in the reference (openai_unetmodel.py:229,685)
            nn.init.zeros_(self.weight)
            nn.init.zeros_(self.bias)
        else:
            fan_in = 1
            for d in wshape[1:]:
                fan_in *= d
            bound = 1.0 / math.sqrt(fan_in)
            nn.init.uniform_(self.weight, -bound, bound)
            if bias:
                nn.init.uniform_(self.bias, -bound, bound)


class _Box(nn.Module):
    """Container whose children are registered under explicit (often numeric) names."""

    def __init__(self, **children):
        super().__init__()
        for k, v in children.items():
            self.add_module(k, v)

    def put(self, name, mod):
        self.add_module(str(name), mod)
        return mod

    def __getitem__(self, i):
        return self._modules[str(i)]


def _res(cin, cout, time_dim):
    b = _Box()
    b.put("in_layers", _Box()).put("0", _Holder((cin,), kind="norm"))
    b.in_layers.put("2", _Holder((cout, cin, 3, 3)))
    b.put("emb_layers", _Box()).put("1", _Holder((cout, time_dim)))
    b.put("out_layers", _Box()).put("0", _Holder((cout,), kind="norm"))
    b.out_layers.put("3", _Holder((cout, cout, 3, 3), kind="zero"))
    if cin != cout:
        b.put("skip_connection", _Holder((cout, cin, 1, 1)))
    return b


def _attn(c, kdim):
    a = _Box()
    a.put("to_q", _Holder((c, c), bias=False))
    a.put("to_k", _Holder((c, kdim), bias=False))
    a.put("to_v", _Holder((c, kdim), bias=False))
    a.put("to_out", _Box()).put("0", _Holder((c, c)))
    return a


def _st(c, context_dim):
    s = _Box()
    s.put("norm", _Holder((c,), kind="norm"))
    s.put("proj_in", _Holder((c, c, 1, 1)))
    t = s.put("transformer_blocks", _Box()).put("0", _Box())
    t.put("attn1", _attn(c, c))
    ff = t.put("ff", _Box()).put("net", _Box())
    ff.put("0", _Box()).put("proj", _Holder((8 * c, c)))
    ff.put("2", _Holder((c, 4 * c)))
    t.put("attn2", _attn(c, context_dim))
    for n in ("norm1", "norm2", "norm3"):
        t.put(n, _Holder((c,), kind="norm"))
    s.put("proj_out", _Holder((c, c, 1, 1), kind="zero"))
    return s


class UNetModelB200(nn.Module):
    def __init__(self, image_size, in_channels, model_channels, out_channels, num_res_blocks,
                 attention_resolutions, dropout=0, channel_mult=(1, 2, 4, 8), conv_resample=True,
                 dims=2, num_classes=None, use_checkpoint=False, use_fp16=False, num_heads=-1,
                 num_head_channels=-1, num_heads_upsample=-1, use_scale_shift_norm=False,
                 resblock_updown=False, use_new_attention_order=False, use_spatial_transformer=False,
                 transformer_depth=1, context_dim=None, n_embed=None, legacy=True,
                 latent_size=(16, 64), max_context_len=40, max_batch=16):
        super().__init__()
        unsupported = []
        if dims != 2: unsupported.append("dims != 2")
        if not conv_resample: unsupported.append("conv_resample=False")
        if num_classes is not None: unsupported.append("num_classes")
        if use_scale_shift_norm: unsupported.append("use_scale_shift_norm")
        if resblock_updown: unsupported.append("resblock_updown")
        if not use_spatial_transformer: unsupported.append("use_spatial_transformer=False")
        if transformer_depth != 1: unsupported.append("transformer_depth != 1")
        if num_heads == -1 or num_head_channels != -1: unsupported.append("num_head_channels")
        if n_embed is not None: unsupported.append("n_embed")
        if dropout: unsupported.append("dropout")
        if context_dim is None: unsupported.append("context_dim=None")
        if unsupported:
            raise NotImplementedError("UNetModelB200 covers the Diff-Foley inference configuration "
                                      "only; unsupported: " + ", ".join(unsupported))
        if isinstance(context_dim, (list, tuple)):
            context_dim = int(context_dim[0])
        self.image_size = image_size
        self.in_channels, self.out_channels = in_channels, out_channels
        self.model_channels = model_channels
        self.num_res_blocks = num_res_blocks
        self.attention_resolutions = tuple(attention_resolutions)
        self.channel_mult = tuple(channel_mult)
        self.num_heads = num_heads
        self.context_dim = context_dim
        self.use_checkpoint = use_checkpoint  # accepted and ignored: inference only
        self.dtype = torch.float32
        self.latent_size = tuple(latent_size)
        self.max_context_len = max_context_len
        self.max_batch = max_batch

        mc, td = model_channels, 4 * model_channels
        self.time_embed = _Box()
        self.time_embed.put("0", _Holder((td, mc)))
        self.time_embed.put("2", _Holder((td, td)))
        # same walk as the reference constructor (openai_unetmodel.py:513-680)
        self.input_blocks = _Box()
        self.input_blocks.put(0, _Box()).put("0", _Holder((mc, in_channels, 3, 3)))
        chans, ch, ds, n_in = [mc], mc, 1, 1
        for level, mult in enumerate(self.channel_mult):
            for _ in range(num_res_blocks):
                blk = self.input_blocks.put(n_in, _Box())
                blk.put("0", _res(ch, mult * mc, td))
                ch = mult * mc
                if ds in self.attention_resolutions:
                    blk.put("1", _st(ch, context_dim))
                chans.append(ch)
                n_in += 1
            if level != len(self.channel_mult) - 1:
                blk = self.input_blocks.put(n_in, _Box())
                blk.put("0", _Box()).put("op", _Holder((ch, ch, 3, 3)))
                chans.append(ch)
                ds *= 2
                n_in += 1
        self.middle_block = _Box()
        self.middle_block.put("0", _res(ch, ch, td))
        self.middle_block.put("1", _st(ch, context_dim))
        self.middle_block.put("2", _res(ch, ch, td))
        self.output_blocks = _Box()
        n_out = 0
        for level, mult in list(enumerate(self.channel_mult))[::-1]:
            for i in range(num_res_blocks + 1):
                ich = chans.pop()
                blk = self.output_blocks.put(n_out, _Box())
                blk.put("0", _res(ch + ich, mc * mult, td))
                ch = mc * mult
                sub = 1
                if ds in self.attention_resolutions:
                    blk.put("1", _st(ch, context_dim))
                    sub = 2
                if level and i == num_res_blocks:
                    blk.put(str(sub), _Box()).put("conv", _Holder((ch, ch, 3, 3)))
                    ds //= 2
                n_out += 1
        self.out = _Box()
        self.out.put("0", _Holder((ch,), kind="norm"))
        self.out.put("2", _Holder((out_channels, mc, 3, 3), kind="zero"))

        self._handle = None
        self._synced = False
        self._device_index = None
        self._fingerprint = None
        # a PARENT module's load_state_dict (LatentDiffusion.init_from_ckpt) recurses through
        # _load_from_state_dict and never calls this module's load_state_dict override: hook the load itself
        self.register_load_state_dict_post_hook(lambda module, incompatible: setattr(module, "_synced", False))

    def _param_fingerprint(self):
        """Cheap change detector for in-place edits (p.copy_, optimizer steps): every in-place op bumps the
        tensor's version counter."""
        return sum(p._version for p in self.parameters()) + 31 * sum(p.data_ptr() % 9973 for p in self.parameters())

    # ------------------------------------------------------------------------------ engine
    def _apply(self, fn, *a, **k):
        self._synced = False
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._synced = False
        return super().load_state_dict(*a, **k)

    def _cfg(self):
        cfg = L.UnetCfg()
        cfg.in_channels, cfg.model_channels = self.in_channels, self.model_channels
        cfg.out_channels, cfg.num_res_blocks = self.out_channels, self.num_res_blocks
        cfg.n_channel_mult = len(self.channel_mult)
        for i, m in enumerate(self.channel_mult):
            cfg.channel_mult[i] = m
        cfg.n_attention_resolutions = len(self.attention_resolutions)
        for i, m in enumerate(self.attention_resolutions):
            cfg.attention_resolutions[i] = m
        cfg.num_heads, cfg.context_dim = self.num_heads, self.context_dim
        cfg.latent_h, cfg.latent_w = self.latent_size
        cfg.max_context_len, cfg.max_batch = self.max_context_len, self.max_batch
        return cfg

    def engine(self, device=None):
        """Creates the C engine on first use and (re)uploads the parameters when they changed."""
        lib = L.lib()
        p0 = next(self.parameters())
        dev = p0.device if device is None else torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("UNetModelB200 runs on a CUDA (sm_100a) device only; move the module "
                               "with .cuda() -- there is no CPU path")
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        if self._handle is not None and self._device_index != idx:
            self.release()
        if self._handle is None:
            h = C.c_void_p()
            cfg = self._cfg()
            L.check(lib.dfb_unet_create(C.byref(cfg), idx, C.byref(h)), "dfb_unet_create")
            self._handle, self._device_index, self._synced = h, idx, False
        if self._synced and self._fingerprint != self._param_fingerprint():
            self._synced = False
        if not self._synced:
            with torch.cuda.device(idx):
                # the pack kernels run on the legacy default stream; conversion temporaries are produced on
                # torch's current stream, which a non-blocking side stream does not order against it
                side = torch.cuda.current_stream(idx) != torch.cuda.default_stream(idx)
                for name, p in self.named_parameters():
                    t = p.detach().to(device=dev, dtype=torch.float32).contiguous()
                    if side:
                        torch.cuda.current_stream(idx).synchronize()
                    shape = (C.c_int64 * t.dim())(*t.shape)
                    L.check(lib.dfb_unet_set_weight(self._handle, name.encode(), L.ptr(t), shape, t.dim()),
                            f"dfb_unet_set_weight({name})")
                    if side:
                        torch.cuda.synchronize(idx)   # `t` may be freed / reused on the side stream next
                torch.cuda.synchronize(idx)
                L.check(lib.dfb_unet_finalize(self._handle), "dfb_unet_finalize")
            self._synced = True
            self._fingerprint = self._param_fingerprint()
        return self._handle

    def release(self):
        if self._handle is not None:
            L.lib().dfb_unet_destroy(self._handle)
            self._handle = None
            self._synced = False

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass

    # ----------------------------------------------------------------------------- forward
    @torch.no_grad()
    def forward(self, x, timesteps=None, context=None, y=None, **kwargs):
        """eps = UNet(x, t, context); x [N,C,H,W] fp32, timesteps [N] int64 or float, context
        [N,L,context_dim] (openai_unetmodel.py:710-742)."""
        assert y is None, "must specify y if and only if the model is class-conditional"
        if timesteps is None or context is None:
            raise ValueError("UNetModelB200.forward needs timesteps and context")
        h = self.engine(x.device)
        n = x.shape[0]
        if tuple(x.shape[1:]) != (self.in_channels, *self.latent_size):
            raise ValueError(f"latent must be [N,{self.in_channels},{self.latent_size[0]},"
                             f"{self.latent_size[1]}], got {tuple(x.shape)}")
        xin = x.detach().to(torch.float32).contiguous()
        ctx = context.detach().to(torch.float32).contiguous()
        if timesteps.dtype.is_floating_point:
            t, t_is_float = timesteps.to(torch.float32).contiguous(), 1
        else:
            t, t_is_float = timesteps.to(torch.int64).contiguous(), 0
        out = torch.empty(n, self.out_channels, *self.latent_size, device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            L.check(L.lib().dfb_unet_forward(h, L.ptr(xin), 1, L.ptr(t), t_is_float, L.ptr(ctx),
                                             ctx.shape[1], L.ptr(out), n, L.cur_stream()),
                    "dfb_unet_forward")
        return out.to(x.dtype)

    @torch.no_grad()
    def profile(self, x, timesteps, context, iters=5):
        """Per-launch event-timed profile of one forward: list of dicts (kind, M, N, K, splits, ctas,
        flops, bytes, ms).  Used by bench.py for the live roofline of the dominant kernel."""
        h = self.engine(x.device)
        n = x.shape[0]
        xin = x.detach().float().contiguous()
        ctx = context.detach().float().contiguous()
        t = timesteps.to(torch.int64).contiguous()
        out = torch.empty(n, self.out_channels, *self.latent_size, device=x.device)
        lib = L.lib()
        with torch.cuda.device(x.device):
            L.check(lib.dfb_unet_set_context(h, L.ptr(ctx), n, ctx.shape[1], L.cur_stream()), "set_context")
            cap = 4096
            infos = (L.OpInfo * cap)()
            n_ops = C.c_int(0)
            L.check(lib.dfb_unet_profile(h, L.ptr(xin), 1, L.ptr(t), 0, L.ptr(out), n, iters, infos, cap,
                                         C.byref(n_ops), L.cur_stream()), "dfb_unet_profile")
        return [dict(kind=infos[i].kind.decode(), M=infos[i].M, N=infos[i].N, K=infos[i].K,
                     splits=infos[i].splits, ctas=infos[i].ctas, flops=infos[i].flops,
                     bytes=infos[i].bytes, ms=infos[i].ms) for i in range(n_ops.value)]

    def trace(self, x, timesteps, context):
        """In-kernel %globaltimer timeline of one graph-replayed forward (diagnostics): int64 array
        [n_ops, 16 marks, 2 (first, last)] in ns, -1 where a mark was not hit (see include/dfb.h)."""
        import numpy as np
        h = self.engine(x.device)
        n = x.shape[0]
        xin = x.detach().float().contiguous()
        ctx = context.detach().float().contiguous()
        t = timesteps.to(torch.int64).contiguous()
        out = torch.empty(n, self.out_channels, *self.latent_size, device=x.device)
        lib = L.lib()
        cap = 4096
        marks = (C.c_ulonglong * (32 * cap))()
        n_ops = C.c_int(0)
        with torch.cuda.device(x.device):
            L.check(lib.dfb_unet_set_context(h, L.ptr(ctx), n, ctx.shape[1], L.cur_stream()), "set_context")
            L.check(lib.dfb_unet_trace(h, L.ptr(xin), 1, L.ptr(t), 0, L.ptr(out), n, marks, cap,
                                       C.byref(n_ops), L.cur_stream()), "dfb_unet_trace")
        a = np.frombuffer(marks, dtype=np.uint64, count=32 * n_ops.value).reshape(n_ops.value, 16, 2).copy()
        hit = a[:, :, 0] != np.uint64(0xFFFFFFFFFFFFFFFF)
        a[:, :, 1] = ~a[:, :, 1]
        res = a.astype(np.int64)
        res[~hit] = -1
        return res

    def debug_taps(self, b_eff):
        """Block outputs of the last forward at this batch size as {name: NCHW tensor} (test aid)."""
        lib, out = L.lib(), {}
        for i in range(lib.dfb_unet_debug_num_taps(self._handle, b_eff)):
            name = C.create_string_buffer(64)
            hwc = (C.c_int32 * 3)()
            L.check(lib.dfb_unet_debug_tap(self._handle, b_eff, i, name, 64, hwc, None, None), "debug_tap")
            t = torch.empty(b_eff, hwc[0], hwc[1], hwc[2], device=f"cuda:{self._device_index}")
            L.check(lib.dfb_unet_debug_tap(self._handle, b_eff, i, name, 64, hwc, L.ptr(t), L.cur_stream()),
                    "debug_tap")
            out[name.value.decode()] = t.permute(0, 3, 1, 2).contiguous()
        torch.cuda.synchronize()
        return out

    def last_launch_count(self):
        return int(L.lib().dfb_unet_last_launch_count(self._handle)) if self._handle else 0
