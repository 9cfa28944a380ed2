"""Multi-GPU DDIM sampling: (clip, guidance-branch) units sharded over the ranks of one box.

The reference's inference path is single-GPU and issues no collectives; this is the
scheme BASELINE.json config 4 names.  Under classifier-free guidance a step evaluates the UNet on
2B independent units (B clips x {uncond, cond}); they share nothing but x_t.  Rank r takes a
contiguous slice of the unit list (ordered like the reference's cat([uncond, cond]) batch,
ddim.py:240-243), runs its UNet forward as one CUDA graph, and the only exchange of the step is an
NCCL all-gather of eps ([units, 4, 16, 64] fp32 = 16 KB per unit) over NVLink; every rank then applies
the fused CFG-combine + DDIM update to all B latents (a few KB of redundant work), so x_t stays
replicated and no second collective is needed.  One process per GPU (torchrun), `torch.distributed`
with the NCCL backend; the same host logic runs under gloo on CPU in the tests with the oracle
standing in for the engine.
"""
import torch
import torch.distributed as dist

from . import _lib as L


def unit_slice(n_clips, world, rank):
    """Units are indexed u = branch * n_clips + clip (branch 0 = uncond, 1 = cond).  Returns the
    (start, stop) slice of rank `rank`; 2*n_clips must be divisible by world."""
    units = 2 * n_clips
    if units % world:
        raise ValueError(f"2*n_clips = {units} units cannot be split evenly over {world} ranks")
    per = units // world
    return rank * per, (rank + 1) * per


def init_library_comm(unet, group=None):
    """Creates the library-owned NCCL communicator of `unet`'s engine (SURVEY 8b 'Ownership'): rank 0
    draws the ncclUniqueId (dfb_comm_unique_id), torch.distributed broadcasts its 128 bytes, every rank
    calls dfb_comm_init.  Idempotent per (engine, world, rank)."""
    import ctypes as C
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = next(unet.parameters()).device
    h = unet.engine(dev)
    key = (world, rank, h.value)
    if getattr(unet, "_comm_key", None) == key:
        return h
    lib = L.lib()
    buf = (C.c_ubyte * 128)()
    if rank == 0:
        L.check(lib.dfb_comm_unique_id(buf), "dfb_comm_unique_id")
    t = torch.tensor(list(buf), dtype=torch.uint8)
    backend = dist.get_backend(group)
    if backend == "nccl":
        t = t.to(dev)
    dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    ident = (C.c_ubyte * 128)(*t.cpu().tolist())
    with torch.cuda.device(dev):
        L.check(lib.dfb_comm_init(h, rank, world, ident), "dfb_comm_init")
    unet._comm_key = key
    return h


@torch.no_grad()
def fused_sharded_ddim_sample(ldm, x_T, cond, uncond, scale, num_steps, group=None):
    """The whole sharded loop as ONE C call per rank (dfb_ddim_sample with a communicator): the per-step
    graph holds this rank's UNet forward over its units, the NCCL all-gather of eps and the fused update of
    all B latents; the embedding table and the step counter are the single-GPU sampler's."""
    import ctypes as C
    from .ddim import DDIMSamplerB200
    unet = ldm.model.diffusion_model
    h = init_library_comm(unet, group)
    sampler = DDIMSamplerB200(ldm)
    sampler.make_schedule(num_steps)
    st = sampler._steps
    dev = x_T.device
    x = x_T.detach().to(torch.float32).clone().contiguous()
    c = cond.detach().to(torch.float32).contiguous()
    u = uncond.detach().to(torch.float32).contiguous()
    fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
    with torch.cuda.device(dev):
        L.check(L.lib().dfb_ddim_sample(
            h, L.ptr(x), L.ptr(c), L.ptr(u), x.shape[0], c.shape[1], float(scale), len(st["timesteps"]),
            st["timesteps"].ctypes.data_as(C.POINTER(C.c_int64)), fp(st["sqrt_one_minus_at"]), fp(st["sqrt_at"]),
            fp(st["sqrt_a_prev"]), fp(st["dir_coef"]), None, None, None, L.cur_stream()), "dfb_ddim_sample(sharded)")
    return x


class _EngineUnits:
    """UNet forward of this rank's units as a replayable CUDA graph around dfb_unet_forward."""

    def __init__(self, ldm, clip_idx, ctx_units):
        from .unet import UNetModelB200
        self.unet = ldm.model.diffusion_model
        if not isinstance(self.unet, UNetModelB200):
            raise RuntimeError("sharded_ddim_sample needs a UNetModelB200 (no fallback path)")
        dev = ctx_units.device
        self.n = ctx_units.shape[0]
        self.clip_idx = clip_idx
        self.h = self.unet.engine(dev)
        self.lib = L.lib()
        self.x_buf = torch.empty(self.n, self.unet.in_channels, *self.unet.latent_size, device=dev)
        self.t_buf = torch.zeros(self.n, dtype=torch.int64, device=dev)
        self.eps = torch.empty_like(self.x_buf)
        with torch.cuda.device(dev):
            self.set_context(ctx_units)
            self._forward()                      # builds the plan (allocations) outside capture
            torch.cuda.synchronize(dev)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._forward()

    def set_context(self, ctx_units):
        """(Re)computes the step-invariant cross-attention K/V of this rank's units."""
        c = ctx_units.float().contiguous()
        L.check(self.lib.dfb_unet_set_context(self.h, L.ptr(c), self.n, c.shape[1], L.cur_stream()),
                "dfb_unet_set_context")

    def _forward(self):
        L.check(self.lib.dfb_unet_forward(self.h, L.ptr(self.x_buf), 1, L.ptr(self.t_buf), 0, None, 0,
                                          L.ptr(self.eps), self.n, L.cur_stream()), "dfb_unet_forward")

    def __call__(self, x, step):
        torch.index_select(x, 0, self.clip_idx, out=self.x_buf)
        self.t_buf.fill_(int(step))
        self.graph.replay()
        return self.eps


@torch.no_grad()
def sharded_ddim_sample(ldm, x_T, cond, uncond, scale, num_steps, group=None, eps_fn=None, step_fn=None):
    """DDIM (eta 0) with CFG for x_T [B,4,H,W], cond/uncond [B,L,D], replicated on every rank.
    Returns the final latents (identical on every rank).

    eps_fn(x_units, t_units, ctx_units) / step_fn(x, e_u, e_c, i) default to the CUDA engine and the
    fused dfb_ddim_step kernel; tests inject the oracle to check the sharding logic on CPU/gloo."""
    from .ddim import DDIMSamplerB200
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if (eps_fn is None and step_fn is None and world > 1 and x_T.is_cuda and dist.get_backend(group) == "nccl"
            and 2 * x_T.shape[0] // world <= ldm.model.diffusion_model.max_batch):
        return fused_sharded_ddim_sample(ldm, x_T, cond, uncond, scale, num_steps, group)
    B = x_T.shape[0]
    lo, hi = unit_slice(B, world, rank)
    dev = x_T.device
    u = torch.arange(lo, hi, device=dev)
    clip_idx = u % B
    ctx_all = torch.cat([uncond, cond])            # unit order, ddim.py:242
    ctx_units = ctx_all.index_select(0, u)
    sampler = DDIMSamplerB200(ldm)
    sampler.make_schedule(num_steps)
    st = sampler._steps
    x = x_T.clone().float()
    eps_all = torch.empty(2 * B, *x.shape[1:], device=dev, dtype=torch.float32)
    engine = None
    if eps_fn is None:
        unet = ldm.model.diffusion_model
        cache = unet.__dict__.setdefault("_units_cache", {})
        key = (lo, hi, B, str(dev), ctx_units.shape[1])
        engine = cache.get(key)
        if engine is None:
            engine = cache[key] = _EngineUnits(ldm, clip_idx, ctx_units)
        else:
            with torch.cuda.device(dev):
                engine.set_context(ctx_units)
    lib = L.lib() if step_fn is None else None
    n = x.numel()
    for i, step in enumerate(st["timesteps"]):
        if engine is not None:
            eps_loc = engine(x, step)
        else:
            eps_loc = eps_fn(x.index_select(0, clip_idx), torch.full((hi - lo,), int(step), dtype=torch.long, device=dev),
                             ctx_units).float().contiguous()
        if world > 1:
            dist.all_gather_into_tensor(eps_all, eps_loc, group=group)
        else:
            eps_all.copy_(eps_loc)
        if step_fn is None:
            with torch.cuda.device(dev):
                L.check(lib.dfb_ddim_step(L.ptr(x), L.ptr(eps_all[:B]), L.ptr(eps_all[B:]), None, float(scale),
                                          float(st["sqrt_one_minus_at"][i]), float(st["sqrt_at"][i]),
                                          float(st["sqrt_a_prev"][i]), float(st["dir_coef"][i]), 0.0, L.ptr(x), None,
                                          n, L.cur_stream()), "dfb_ddim_step")
        else:
            x = step_fn(x, eps_all[:B], eps_all[B:], i)
    return x
