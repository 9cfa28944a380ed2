"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by the product path (diff_foley_b200/).

CPU restatement of the reference's DDIM sampler arithmetic and noise schedule, plus the
cond-stage embedder.  Cites luosiallen/Diff-Foley @ 0ba1e8ad:
  diff_foley/models/diffusion/ddpm.py:122-174        register_schedule
  diff_foley/modules/diffusionmodules/util.py:21-74   make_beta_schedule / make_ddim_timesteps /
                                                      make_ddim_sampling_parameters
  diff_foley/models/diffusion/ddim.py:27-56,179-273   make_schedule / ddim_sampling / p_sample_ddim
  diff_foley/modules/cond_stage/video_feat_encoder.py:12-18
Pinned by tests/golden/make_golden.py against the reference's own DDIMSampler (CPU).
"""
import numpy as np
import torch
import torch.nn.functional as F

# inference/config/Stage2_LDM.yaml:5-9
LINEAR_START, LINEAR_END, NUM_TIMESTEPS = 0.00085, 0.0120, 1000


def alphas_cumprod(linear_start=LINEAR_START, linear_end=LINEAR_END, n=NUM_TIMESTEPS):
    """'linear' schedule (util.py:22-25) -> cumulative product (ddpm.py:129-130), stored as fp32
    exactly like register_buffer(to_torch(...)) does (ddpm.py:139-142)."""
    betas = torch.linspace(linear_start ** 0.5, linear_end ** 0.5, n, dtype=torch.float64) ** 2
    ac = np.cumprod(1.0 - betas.numpy(), axis=0)
    return torch.tensor(ac, dtype=torch.float32)


def ddim_timesteps(num_ddim_steps, num_ddpm_steps=NUM_TIMESTEPS):
    """'uniform' discretisation with the reference's +1 offset (util.py:46-60)."""
    c = num_ddpm_steps // num_ddim_steps
    return np.asarray(list(range(0, num_ddpm_steps, c))) + 1


def ddim_coefficients(num_ddim_steps, eta=0.0, ac=None):
    """Per-step fp32 scalars in SAMPLING order (t descending), computed with the same fp32
    operation order the reference ends up with (ddim.py:46-52 buffers, :251-270 use):
      a_t = ac[t], a_prev = ac[t_prev] (ac[0] for the last step)     util.py:65-66
      sqrt_one_minus_at = sqrt(1 - a_t)        fp32 tensor op         ddim.py:52
      sqrt_at = a_t.sqrt()                     fp32                   ddim.py:258
      sqrt_a_prev = a_prev.sqrt()              fp32                   ddim.py:272
      dir = (1 - a_prev - sigma_t**2).sqrt()   fp32, sigma = 0 @eta 0 ddim.py:263
    Returns dict of numpy arrays: timesteps int64[S], and fp32[S] coefficient arrays."""
    assert eta == 0.0, "the hot path fixes eta = 0 (deterministic DDIM)"
    if ac is None:
        ac = alphas_cumprod()
    ts = ddim_timesteps(num_ddim_steps, ac.shape[0])
    a = ac[ts]                                                       # fp32 tensor
    a_prev = torch.tensor([ac[0].item()] + ac[ts[:-1]].tolist(), dtype=torch.float32)
    sqrt_1m = torch.sqrt(1.0 - a)
    one = torch.tensor(1.0, dtype=torch.float32)
    sigma = torch.zeros_like(a)
    dirc = (one - a_prev - sigma ** 2).sqrt()
    order = np.arange(len(ts))[::-1].copy()                          # np.flip(timesteps), ddim.py:199
    return dict(
        timesteps=ts[order].astype(np.int64),
        sqrt_one_minus_at=sqrt_1m.numpy()[order].copy(),
        sqrt_at=a.sqrt().numpy()[order].copy(),
        sqrt_a_prev=a_prev.sqrt().numpy()[order].copy(),
        dir_coef=dirc.numpy()[order].copy(),
        a_t=a.numpy()[order].copy(),
    )


def ddim_step(x, e_uncond, e_cond, scale, c, i, grad=None):
    """One p_sample_ddim update (ddim.py:241-245 CFG, :377-380 classifier term, :258-273), fp32,
    same operation order.  c = ddim_coefficients(...), i = index in sampling order."""
    f = lambda v: torch.tensor(float(v), dtype=torch.float32)
    e = e_uncond + f(scale) * (e_cond - e_uncond) if e_uncond is not None else e_cond
    if grad is not None:
        e = e - (f(1.0) - f(c["a_t"][i])).sqrt() * grad
    pred_x0 = (x - f(c["sqrt_one_minus_at"][i]) * e) / f(c["sqrt_at"][i])
    x_prev = f(c["sqrt_a_prev"][i]) * pred_x0 + f(c["dir_coef"][i]) * e
    return x_prev, pred_x0


@torch.no_grad()
def ddim_sample(eps_fn, x_T, cond, uncond, scale, num_steps):
    """ddim_sampling loop (ddim.py:204-228) with classifier-free guidance batching (:240-244):
    eps_fn(x_in [2B,...], t_in int64 [2B], c_in [2B,L,D]) -> eps [2B,...]."""
    c = ddim_coefficients(num_steps)
    x = x_T.clone()
    b = x.shape[0]
    pred = None
    for i, step in enumerate(c["timesteps"]):
        ts = torch.full((b,), int(step), dtype=torch.long)
        out = eps_fn(torch.cat([x, x]), torch.cat([ts, ts]), torch.cat([uncond, cond]))
        e_u, e_c = out.chunk(2)
        x, pred = ddim_step(x, e_u, e_c, scale, c, i)
    return x, pred


def cond_stage(sd, feats):
    """Video_Feat_Encoder_Posembed.forward (video_feat_encoder.py:12-18): Linear(512->768) +
    learned positional embedding of the first seq_len rows."""
    x = F.linear(feats, sd["embedder.0.weight"], sd["embedder.0.bias"])
    return x + sd["pos_emb.weight"][: feats.shape[1]][None]


def cond_stage_seeded_state(seed, origin_dim=512, embed_dim=768, seq_len=40):
    """Seeded parameters of the cond-stage embedder under the reference's keys (regenerated identically by the
    golden generator and by the tests, so the 1.6 MB of weights never enter a fixture)."""
    g = torch.Generator().manual_seed(seed)
    shapes = {"embedder.0.weight": (embed_dim, origin_dim), "embedder.0.bias": (embed_dim,),
              "pos_emb.weight": (seq_len, embed_dim)}
    return {k: torch.randn(v, generator=g) * (0.05 if len(v) > 1 else 0.1) for k, v in shapes.items()}
