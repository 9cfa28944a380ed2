"""Multi-GPU parity check (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/check_sharded.py
The fused sharded sampler (library-owned NCCL communicator, eps all-gather inside the step graph) must return
the same latents on every rank, and they must equal the single-GPU fused sampler's bit for bit when the
per-rank unit count matches the single-GPU batch (same plan, same kernels), else to fp16-rounding noise."""
import os
import sys

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import numpy as np
import torch
import torch.distributed as dist

from diff_foley_b200 import _lib as L
from diff_foley_b200.ldm import LatentDiffusionB200
from diff_foley_b200.parallel import sharded_ddim_sample
from diff_foley_b200.unet import UNetModelB200
from helpers import unet_kwargs
from oracle import unet_oracle

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
cfg = unet_oracle.small_unet_cfg()
g = np.load("tests/golden/ddim_small.npz")
unet = UNetModelB200(**unet_kwargs(cfg), max_batch=16)
unet.load_state_dict(unet_oracle.seeded_state_dict(cfg, int(g["seed"])))
unet = unet.to(dev)
ldm = LatentDiffusionB200(unet, cond_stage_params=dict(origin_dim=64, embed_dim=cfg["context_dim"], seq_len=40)).to(dev)
gen = torch.Generator().manual_seed(11)
B = 2 * world
x_T = torch.randn(B, 4, cfg["latent_h"], cfg["latent_w"], generator=gen).to(dev)
cond = torch.randn(B, cfg["context_len"], cfg["context_dim"], generator=gen).to(dev)
unc = torch.zeros_like(cond)
out = sharded_ddim_sample(ldm, x_T, cond, unc, 4.5, 25)
torch.cuda.synchronize()
gathered = [torch.empty_like(out) for _ in range(world)]
dist.all_gather(gathered, out)
same = all(torch.equal(gathered[0], t) for t in gathered)
# single-GPU fused sampler on the same inputs (communicator dropped)
L.check(L.lib().dfb_comm_destroy(unet.engine(dev)), "dfb_comm_destroy")
unet._comm_key = None
ref, _ = ldm.sample_log_diff_sampler(cond, B, "DDIM", 25, size_len=cfg["latent_w"], unconditional_guidance_scale=4.5,
                                     unconditional_conditioning=unc, x_T=x_T)
torch.cuda.synchronize()
err = float((out.double() - ref.double()).norm() / ref.double().norm())
# and the reference sampler's golden (first 2 clips of ddim_small use its own x_T / cond)
x2, c2 = torch.from_numpy(g["x_T"]).to(dev), torch.from_numpy(g["cond"]).to(dev)
if (2 * x2.shape[0]) % world == 0:
    unet._comm_key = None
    o2 = sharded_ddim_sample(ldm, x2, c2, torch.zeros_like(c2), float(g["scale"]), int(g["steps"]))
    torch.cuda.synchronize()
    e2 = float((o2.cpu().double() - torch.from_numpy(g["samples"]).double()).norm() / torch.from_numpy(g["samples"]).double().norm())
else:
    e2 = float("nan")
if rank == 0:
    print(f"world {world}: identical on all ranks = {same}; sharded vs single-GPU fused rel-L2 = {err:.3e}; "
          f"sharded vs reference golden (ddim_small) rel-L2 = {e2:.3e}")
    assert same and err < 1e-3 and (e2 != e2 or e2 < 1e-3)
dist.barrier()
dist.destroy_process_group()
