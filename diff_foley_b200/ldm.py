"""Host-side mirror of the parts of LatentDiffusion the sampling hot path touches
(diff_foley/models/diffusion/ddpm.py: register_schedule :122-174, get_learned_conditioning :568-579,
apply_model :925-1026, sample_log_diff_sampler :1287-1314, DiffusionWrapper :1545-1571) and of the
cond-stage embedder (diff_foley/modules/cond_stage/video_feat_encoder.py:4-18).

In a checkout that has the reference installed you keep the reference's own LatentDiffusion and only
swap `unet_config.target` to diff_foley_b200.unet.UNetModelB200.  This module is
the same surface for hosts without the reference (the GPU box, bench.py, the tests): same attribute
names (`model.diffusion_model`, `cond_stage_model`, `alphas_cumprod`, ...), same state-dict keys
(`model.diffusion_model.*`, `cond_stage_model.*`), same method signatures.
"""
import numpy as np
import torch
import torch.nn as nn

from . import _lib as L
from .ddim import DDIMSamplerB200
from .unet import UNetModelB200


class VideoFeatEncoderPosembed(nn.Module):
    """Linear(origin_dim -> embed_dim) + learned positional embedding (video_feat_encoder.py:4-18).
    Parameter names match the reference (`embedder.0.*`, `pos_emb.weight`).  The projection runs on
    the tcgen05 GEMM (`dfb_gemm`, bias + positional rows fused as the residual epilogue)."""

    def __init__(self, origin_dim, embed_dim, seq_len=215):
        super().__init__()
        self.embedder = nn.Sequential(nn.Linear(origin_dim, embed_dim))
        self.pos_emb = nn.Embedding(seq_len, embed_dim)

    @torch.no_grad()
    def forward(self, x):
        bs, seq_len, c = x.shape
        if x.device.type != "cuda":
            raise RuntimeError("VideoFeatEncoderPosembed runs on a CUDA (sm_100a) device only")
        lin = self.embedder[0]
        a16 = x.reshape(bs * seq_len, c).to(torch.float16).contiguous()
        w16 = lin.weight.detach().to(torch.float16).contiguous()
        bias = lin.bias.detach().float().contiguous()
        pos = self.pos_emb.weight.detach()[:seq_len].float().repeat(bs, 1).contiguous()
        out = torch.empty(bs * seq_len, lin.out_features, device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            L.check(L.lib().dfb_gemm(L.ptr(a16), L.ptr(w16), bs * seq_len, lin.out_features, c, L.ptr(bias),
                                     L.ptr(pos), 0, L.ptr(out), None, 1, L.cur_stream()), "dfb_gemm(cond_stage)")
        return out.view(bs, seq_len, lin.out_features)


class DiffusionWrapperB200(nn.Module):
    def __init__(self, unet, conditioning_key="crossattn"):
        super().__init__()
        assert conditioning_key == "crossattn", "Diff-Foley conditions through cross-attention only"
        self.diffusion_model = unet
        self.conditioning_key = conditioning_key

    def forward(self, x, t, c_concat=None, c_crossattn=None):
        cc = torch.cat(c_crossattn, 1)                      # ddpm.py:1559-1560
        return self.diffusion_model(x, t, context=cc)


class LatentDiffusionB200(nn.Module):
    def __init__(self, unet_params, cond_stage_params=None, linear_start=0.00085, linear_end=0.0120,
                 timesteps=1000, channels=4, scale_factor=0.18215, conditioning_key="crossattn",
                 first_stage_params=None, **ignored):
        super().__init__()
        unet = unet_params if isinstance(unet_params, nn.Module) else UNetModelB200(**unet_params)
        self.model = DiffusionWrapperB200(unet, conditioning_key)
        cs = dict(origin_dim=512, embed_dim=768, seq_len=40) if cond_stage_params is None else cond_stage_params
        self.cond_stage_model = VideoFeatEncoderPosembed(**cs)
        # first stage (decode half only): built when asked for -- `first_stage_params` = the reference's
        # first_stage_config.params (Stage2_LDM.yaml:38-59) or True for its defaults
        self.first_stage_model = None
        if first_stage_params:
            from .vae import AutoencoderKLDecoderB200
            fs = {} if first_stage_params is True else dict(first_stage_params)
            self.first_stage_model = AutoencoderKLDecoderB200(ddconfig=fs.get("ddconfig"),
                                                              embed_dim=fs.get("embed_dim", 4),
                                                              scale_factor=scale_factor)
        self.channels = channels
        self.scale_factor = scale_factor
        self.parameterization = "eps"
        # 'linear' beta schedule (util.py:22-25) and its cumulative products, stored as fp32 buffers
        betas = (torch.linspace(linear_start ** 0.5, linear_end ** 0.5, timesteps, dtype=torch.float64) ** 2).numpy()
        ac = np.cumprod(1. - betas, axis=0)
        self.num_timesteps = int(timesteps)
        f32 = lambda a: torch.tensor(a, dtype=torch.float32)
        self.register_buffer("betas", f32(betas))
        self.register_buffer("alphas_cumprod", f32(ac))
        self.register_buffer("alphas_cumprod_prev", f32(np.append(1., ac[:-1])))

    @property
    def device(self):
        return self.betas.device

    def get_learned_conditioning(self, c):
        return self.cond_stage_model(c)

    @torch.no_grad()
    def decode_first_stage(self, z, predict_cids=False, force_not_quantize=False):
        """ddpm.py:739-797: `1/scale_factor * z` -> AutoencoderKL.decode -> [B,3,128,512] (mel = channel 0)."""
        if self.first_stage_model is None:
            raise RuntimeError("LatentDiffusionB200 was built without first_stage_params")
        return self.first_stage_model.decode(z / self.scale_factor)

    def apply_model(self, x_noisy, t, cond, return_ids=False):
        if not isinstance(cond, list):
            cond = [cond]
        return self.model(x_noisy, t, c_crossattn=cond)

    @torch.no_grad()
    def sample_log_diff_sampler(self, cond, batch_size, sampler_name, ddim_steps, size_len=64,
                                unconditional_guidance_scale=1.0, unconditional_conditioning=None, **kwargs):
        shape = (self.channels, 16, size_len)                # ddpm.py:1293
        if sampler_name == "DDIM":
            sampler = DDIMSamplerB200(self)
        elif sampler_name == "DPM_Solver":                   # ddpm.py:1297-1301 (the notebook's default)
            from .dpm_solver import DPMSolverSamplerB200
            sampler = DPMSolverSamplerB200(self)
        elif sampler_name == "PLMS":                         # ddpm.py:1303-1307
            from .plms import PLMSSamplerB200
            sampler = PLMSSamplerB200(self)
        else:
            raise NotImplementedError(f"sampler_name {sampler_name!r}: DDIM, DPM_Solver and PLMS are implemented "
                                      "(the reference's fall-through is its ancestral DDPM loop, outside the hot path)")
        return sampler.sample(ddim_steps, batch_size, shape, cond, verbose=False,
                              unconditional_guidance_scale=unconditional_guidance_scale,
                              unconditional_conditioning=unconditional_conditioning, **kwargs)

    @torch.no_grad()
    def sample_log_with_classifier_diff_sampler(self, embed_cond, origin_cond, batch_size, sampler_name="DDIM",
                                                ddim_steps=250, size_len=64, unconditional_guidance_scale=1.0,
                                                unconditional_conditioning=None, classifier=None,
                                                classifier_guide_scale=0.0, **kwargs):
        shape = (self.channels, 16, size_len)                # ddpm.py:1341
        if sampler_name == "DPM_Solver":                     # ddpm.py:1346-1350
            from .dpm_solver import DPMSolverSamplerB200
            return DPMSolverSamplerB200(self).sample_with_classifier(
                ddim_steps, batch_size, shape, embed_cond, origin_cond=origin_cond,
                unconditional_guidance_scale=unconditional_guidance_scale,
                unconditional_conditioning=unconditional_conditioning, classifier=classifier,
                classifier_guide_scale=classifier_guide_scale, **kwargs)
        if sampler_name != "DDIM":
            raise NotImplementedError(f"sampler_name {sampler_name!r}: DDIM and DPM_Solver take classifier guidance")
        return DDIMSamplerB200(self).sample_with_classifier(
            ddim_steps, batch_size, shape, embed_cond, origin_cond=origin_cond, verbose=False,
            unconditional_guidance_scale=unconditional_guidance_scale,
            unconditional_conditioning=unconditional_conditioning, classifier=classifier,
            classifier_guide_scale=classifier_guide_scale, **kwargs)
