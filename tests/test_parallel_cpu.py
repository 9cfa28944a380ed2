"""world_size-2 gloo test of the multi-GPU sharding logic (unit partition, all-gather order, replicated
update) with the oracle standing in for the CUDA engine: the sharded sampler must reproduce the
single-process oracle sampler exactly."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diff_foley_b200.parallel import sharded_ddim_sample, unit_slice
from oracle import ddim_oracle, unet_oracle

CFG = unet_oracle.small_unet_cfg(model_channels=64, channel_mult=(1, 2), num_heads=4, context_dim=64,
                                 latent_h=8, latent_w=16, context_len=8, attention_resolutions=(2, 1))


class _LDM:  # what DDIMSamplerB200.make_schedule reads
    num_timesteps = 1000
    alphas_cumprod = ddim_oracle.alphas_cumprod()


def _inputs():
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 4, CFG["latent_h"], CFG["latent_w"], generator=g)
    cond = torch.randn(2, CFG["context_len"], CFG["context_dim"], generator=g)
    return x, cond, torch.zeros_like(cond)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    sd = unet_oracle.seeded_state_dict(CFG, 11)
    x, cond, unc = _inputs()
    c = ddim_oracle.ddim_coefficients(5)
    res = sharded_ddim_sample(
        _LDM(), x, cond, unc, 4.5, 5,
        eps_fn=lambda a, t, ctx: unet_oracle.unet_forward(sd, CFG, a, t, ctx),
        step_fn=lambda xx, eu, ec, i: ddim_oracle.ddim_step(xx, eu, ec, 4.5, c, i)[0])
    out[rank] = res.clone()
    dist.barrier()
    dist.destroy_process_group()


def test_unit_slice():
    assert unit_slice(1, 2, 0) == (0, 1) and unit_slice(1, 2, 1) == (1, 2)   # cond / uncond of one clip
    assert unit_slice(64, 8, 3) == (48, 64)                                  # BASELINE config 4: 16 units / GPU
    with pytest.raises(ValueError):
        unit_slice(3, 4, 0)


def test_sharded_sampler_world2_matches_single_process():
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    sd = unet_oracle.seeded_state_dict(CFG, 11)
    x, cond, unc = _inputs()
    want, _ = ddim_oracle.ddim_sample(lambda a, t, c: unet_oracle.unet_forward(sd, CFG, a, t, c), x, cond, unc, 4.5, 5)
    assert torch.equal(out[0], out[1])              # replicated latents
    # batch-of-2 vs two batches-of-1 through ATen: equal up to fp32 summation order
    assert float((out[0] - want).norm() / want.norm()) < 1e-5
