"""DDIMSamplerB200 -- the reference's DDIMSampler interface (diff_foley/models/diffusion/ddim.py)
on top of the fused CUDA sampler in libdfb.so.

`DDIMSamplerB200(model).sample(S, batch_size, shape, conditioning, eta=0., x_T=...,
unconditional_guidance_scale=..., unconditional_conditioning=...) -> (samples, intermediates)` has
the signature and return value of ddim.py:58-113.  With classifier-free guidance on a
UNetModelB200 the whole 25-step loop is ONE C call (`dfb_ddim_sample`): the step (UNet forward on
the cond/uncond pair + CFG combine + DDIM update) is captured once as a CUDA graph and replayed,
cross-attention K/V of the step-invariant context are computed once per clip.  Other call shapes
(no guidance, classifier guidance, callbacks) run the host loop of ddim.py:204-228 with
`model.apply_model` + the fused `dfb_ddim_step` kernel.  Options the hot path does not cover
(eta > 0, mask/x0 inpainting, quantize_denoised, score_corrector) raise NotImplementedError.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L
from .unet import UNetModelB200


def make_ddim_timesteps(num_ddim_timesteps, num_ddpm_timesteps):
    """'uniform' discretisation with the +1 offset (util.py:46-60): 1, 41, ..., 961 for S = 25."""
    c = num_ddpm_timesteps // num_ddim_timesteps
    return np.asarray(list(range(0, num_ddpm_timesteps, c))) + 1


class DDIMSamplerB200(object):
    def __init__(self, model, schedule="linear", **kwargs):
        self.model = model
        self.ddpm_num_timesteps = model.num_timesteps
        self.schedule = schedule

    # ------------------------------------------------------------------------------ schedule
    def make_schedule(self, ddim_num_steps, ddim_discretize="uniform", ddim_eta=0., verbose=False):
        """fp32 per-step scalars with the reference's operation order (ddim.py:27-56, 251-270)."""
        if ddim_discretize != "uniform":
            raise NotImplementedError("only the 'uniform' DDIM discretisation is supported")
        if ddim_eta != 0.:
            raise NotImplementedError("the B200 sampler implements deterministic DDIM (eta = 0)")
        ac = self.model.alphas_cumprod.detach().to("cpu", torch.float32)
        assert ac.shape[0] == self.ddpm_num_timesteps
        ts = make_ddim_timesteps(ddim_num_steps, self.ddpm_num_timesteps)
        a = ac[ts]
        a_prev = torch.tensor([ac[0].item()] + ac[ts[:-1]].tolist(), dtype=torch.float32)
        self.ddim_timesteps = ts
        self.ddim_alphas = a
        self.ddim_alphas_prev = a_prev
        self.ddim_sigmas = torch.zeros_like(a)
        self.ddim_sqrt_one_minus_alphas = torch.sqrt(1. - a)
        order = np.arange(len(ts))[::-1].copy()
        one = torch.tensor(1.0, dtype=torch.float32)
        self._steps = dict(
            timesteps=np.ascontiguousarray(ts[order].astype(np.int64)),
            sqrt_one_minus_at=np.ascontiguousarray(self.ddim_sqrt_one_minus_alphas.numpy()[order]),
            sqrt_at=np.ascontiguousarray(a.sqrt().numpy()[order]),
            sqrt_a_prev=np.ascontiguousarray(a_prev.sqrt().numpy()[order]),
            dir_coef=np.ascontiguousarray((one - a_prev - self.ddim_sigmas ** 2).sqrt().numpy()[order]),
            grad_coef=np.ascontiguousarray((one - a).sqrt().numpy()[order]),
        )

    # ------------------------------------------------------------------------------- sample
    def _unet(self):
        m = getattr(getattr(self.model, "model", None), "diffusion_model", None)
        return m if isinstance(m, UNetModelB200) else None

    @torch.no_grad()
    def sample(self, S, batch_size, shape, conditioning=None, callback=None, normals_sequence=None,
               img_callback=None, quantize_x0=False, eta=0., mask=None, x0=None, temperature=1.,
               noise_dropout=0., score_corrector=None, corrector_kwargs=None, verbose=True, x_T=None,
               log_every_t=100, unconditional_guidance_scale=1., unconditional_conditioning=None,
               origin_cond=None, classifier=None, classifier_guide_scale=0.0, **kwargs):
        if quantize_x0 or mask is not None or x0 is not None or score_corrector is not None:
            raise NotImplementedError("quantize_x0 / mask / score_corrector are outside the hot path")
        self.make_schedule(ddim_num_steps=S, ddim_eta=eta, verbose=verbose)
        st = self._steps
        device = self.model.betas.device
        C_, H, W = shape
        if x_T is None:
            img = torch.randn((batch_size, C_, H, W), device=device)
        else:
            img = x_T.to(device=device, dtype=torch.float32).clone()
        intermediates = {"x_inter": [img.clone()], "pred_x0": [img.clone()]}
        cfg = not (unconditional_conditioning is None or unconditional_guidance_scale == 1.)
        unet = self._unet()
        fused = (cfg and unet is not None and classifier is None and callback is None and
                 img_callback is None and 2 * batch_size <= unet.max_batch)
        lib = L.lib()
        if fused:
            cond = conditioning.to(device=device, dtype=torch.float32).contiguous()
            unc = unconditional_conditioning.to(device=device, dtype=torch.float32).contiguous()
            # raw pointers cross the C ABI: every extent the library derives from its config is checked here
            if (C_, H, W) != (unet.in_channels, *unet.latent_size):
                raise ValueError(f"shape {tuple(shape)} does not match the UNet latent "
                                 f"({unet.in_channels}, {unet.latent_size[0]}, {unet.latent_size[1]})")
            if tuple(img.shape) != (batch_size, C_, H, W):
                raise ValueError(f"x_T must be [{batch_size},{C_},{H},{W}], got {tuple(img.shape)}")
            if (cond.dim() != 3 or cond.shape != unc.shape or cond.shape[0] != batch_size or
                    cond.shape[1] > unet.max_context_len or cond.shape[2] != unet.context_dim):
                raise ValueError(f"conditioning / unconditional_conditioning must both be [{batch_size}, L <= "
                                 f"{unet.max_context_len}, {unet.context_dim}], got {tuple(cond.shape)} / {tuple(unc.shape)}")
            h = unet.engine(device)
            n_steps = len(st["timesteps"])      # > S when S does not divide 1000 (ddim.py:197-199 runs them all)
            pred = torch.empty_like(img)
            x1, p1 = torch.empty_like(img), torch.empty_like(img)
            fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
            with torch.cuda.device(device):
                L.check(lib.dfb_ddim_sample(
                    h, L.ptr(img), L.ptr(cond), L.ptr(unc), batch_size, cond.shape[1],
                    float(unconditional_guidance_scale), n_steps,
                    st["timesteps"].ctypes.data_as(C.POINTER(C.c_int64)), fp(st["sqrt_one_minus_at"]),
                    fp(st["sqrt_at"]), fp(st["sqrt_a_prev"]), fp(st["dir_coef"]), L.ptr(pred), L.ptr(x1), L.ptr(p1),
                    L.cur_stream()), "dfb_ddim_sample")
            # the reference logs after the step with index == total_steps - 1 (the first) and after every step
            # with index % log_every_t == 0 (ddim.py:223-226) -- always the last one (index 0); the fused loop
            # exposes exactly those two, i.e. the reference's list whenever total_steps <= log_every_t
            if n_steps > 1:
                intermediates["x_inter"].append(x1)
                intermediates["pred_x0"].append(p1)
            intermediates["x_inter"].append(img.clone())
            intermediates["pred_x0"].append(pred)
            return img, intermediates
        # ---- host loop (ddim.py:204-228) for the call shapes the fused sampler does not take
        n = img.numel()
        pred = torch.empty_like(img)
        total_steps = len(st["timesteps"])
        for i, step in enumerate(st["timesteps"]):
            index = total_steps - i - 1
            ts = torch.full((batch_size,), int(step), device=device, dtype=torch.long)
            if cfg:
                e = self.model.apply_model(torch.cat([img] * 2), torch.cat([ts] * 2),
                                           torch.cat([unconditional_conditioning, conditioning]))
                e = e.to(torch.float32).contiguous()
                e_u, e_c = e[:batch_size], e[batch_size:]
            else:
                e_u, e_c = None, self.model.apply_model(img, ts, conditioning).to(torch.float32).contiguous()
            grad = None
            if classifier is not None:  # ddim.py:333-341
                if hasattr(classifier, "loglikelihood_grad"):      # native forward + backward (classifier.py)
                    grad = classifier.loglikelihood_grad(img, ts, origin_cond, classifier_guide_scale)
                else:                                              # a foreign classifier module: its own autograd
                    with torch.enable_grad():
                        x_in = img.detach().requires_grad_(True)
                        log_probs = torch.log(classifier(x_in, t=ts, video_feat=origin_cond))
                        grad = (torch.autograd.grad(log_probs.sum(), x_in)[0] * classifier_guide_scale)
                grad = grad.to(torch.float32).contiguous()
            nxt = torch.empty_like(img)
            with torch.cuda.device(device):
                L.check(lib.dfb_ddim_step(
                    L.ptr(img), L.ptr(e_u), L.ptr(e_c), L.ptr(grad), float(unconditional_guidance_scale),
                    float(st["sqrt_one_minus_at"][i]), float(st["sqrt_at"][i]), float(st["sqrt_a_prev"][i]),
                    float(st["dir_coef"][i]), float(st["grad_coef"][i]), L.ptr(nxt), L.ptr(pred), n,
                    L.cur_stream()), "dfb_ddim_step")
            img = nxt
            if callback: callback(i)
            if img_callback: img_callback(pred, i)
            if index % log_every_t == 0 or index == total_steps - 1:
                intermediates["x_inter"].append(img)
                intermediates["pred_x0"].append(pred.clone())
        return img, intermediates

    @torch.no_grad()
    def sample_with_classifier(self, S, batch_size, shape, conditioning=None, origin_cond=None, **kw):
        """ddim.py:115-176; classifier / classifier_guide_scale come in through **kw."""
        return self.sample(S, batch_size, shape, conditioning=conditioning, origin_cond=origin_cond, **kw)
