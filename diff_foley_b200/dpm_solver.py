"""DPMSolverSamplerB200 -- the reference's DPMSolverSampler interface
(diff_foley/models/diffusion/dpm_solver/sampler.py:11-156; the notebook's default sampler) on the fused
CUDA sampler of libdfb.so.

The reference wires NoiseScheduleVP('discrete') + model_wrapper('classifier-free') + DPM_Solver(predict_x0=True)
.sample(steps=S, skip_type='time_uniform', method='multistep', order=2, lower_order_final=True)
(sampler.py:67-85).  For that fixed configuration every per-step scalar -- the fractional model-input time, the
marginal mean / std of the data prediction, and the three coefficients of the first- / second-order multistep
update (dpm_solver.py:504-533, 755-790) -- depends on the schedule only, so the host computes them once in fp32
and the whole loop is ONE C call (`dfb_dpm_solver_sample`): the step (UNet forward on the cond / uncond pair at
a fractional time served from the per-schedule embedding table + CFG combine + data prediction + multistep
update) is one CUDA graph replayed S times.

`sample_with_classifier` (double guidance, sampler.py:89-156, dpm_solver.py:1352-1393) runs the same updates
as a host loop, because the classifier gradient enters between the UNet and the update.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L
from .unet import UNetModelB200


class _NoiseScheduleDiscrete:
    """What NoiseScheduleVP('discrete', alphas_cumprod=...) evaluates (dpm_solver.py:98-156), fp32."""

    def __init__(self, alphas_cumprod):
        self.log_alpha = (0.5 * torch.log(alphas_cumprod.detach().to("cpu", torch.float32))).contiguous()
        self.total_N = int(self.log_alpha.shape[0])
        self.t_array = torch.linspace(0., 1., self.total_N + 1)[1:]

    def log_mean_coeff(self, t):
        """piecewise-linear in t through (t_array, log_alpha), end segments extended (interpolate_fn :1132-1171)"""
        xp, yp = self.t_array, self.log_alpha
        j = (torch.searchsorted(xp, t, right=False) - 1).clamp(0, self.total_N - 2)
        return yp[j] + (t - xp[j]) * (yp[j + 1] - yp[j]) / (xp[j + 1] - xp[j])

    def alpha(self, t):
        return torch.exp(self.log_mean_coeff(t))

    def std(self, t):
        return torch.sqrt(1. - torch.exp(2. * self.log_mean_coeff(t)))

    def lam(self, t):
        lm = self.log_mean_coeff(t)
        return lm - 0.5 * torch.log(1. - torch.exp(2. * lm))

    def model_input_time(self, t):
        return (t - 1. / self.total_N) * 1000.


def dpm_solver_pp_2m_schedule(alphas_cumprod, steps, order=2, lower_order_final=True):
    """Per-evaluation scalars of multistep DPM-Solver++ (fp32 numpy arrays of `steps` entries, k = 0..S-1:
    the model is evaluated at t_k and the update takes x from t_k to t_{k+1}), see include/dfb.h."""
    if steps < order:
        raise ValueError("DPM-Solver needs steps >= order")
    ns = _NoiseScheduleDiscrete(alphas_cumprod)
    t = torch.linspace(1.0, 1. / ns.total_N, steps + 1)                     # time_uniform, t_T = 1, t_0 = 1/N
    lam, sig, alp = ns.lam(t), ns.std(t), ns.alpha(t)
    k = torch.arange(steps)
    h = lam[k + 1] - lam[k]
    cx = sig[k + 1] / sig[k]
    a2 = alp[k + 1] * (torch.exp(-h) - 1.)                                  # second-order form (:771-775)
    a1 = alp[k + 1] * torch.expm1(-h)                                       # first-order form (:524-531)
    orders = torch.full((steps,), order, dtype=torch.int32)
    orders[0] = 1                                                           # init by the lower-order solver (:1073)
    if lower_order_final and steps < 15:                                    # :1081-1084
        orders[steps - 1] = min(order, 1)
    inv_r0 = torch.zeros(steps)
    inv_r0[1:] = 1. / ((lam[1:steps] - lam[0:steps - 1]) / h[1:])           # 1 / r0, r0 = h_0 / h (:766-768)
    a = torch.where(orders == 1, a1, a2)
    f = lambda v: np.ascontiguousarray(v.to(torch.float32).numpy())
    return dict(t_cont=f(t), t_input=f(ns.model_input_time(t[:steps])), sigma=f(sig[:steps]), alpha=f(alp[:steps]),
                cx=f(cx), a_coef=f(a), inv_r0=f(inv_r0), order=np.ascontiguousarray(orders.numpy().astype(np.int32)))


class DPMSolverSamplerB200(object):
    def __init__(self, model, **kwargs):
        self.model = model
        self.alphas_cumprod = model.alphas_cumprod.detach().to(torch.float32)

    def _unet(self):
        m = getattr(getattr(self.model, "model", None), "diffusion_model", None)
        return m if isinstance(m, UNetModelB200) else None

    @torch.no_grad()
    def sample(self, S, batch_size, shape, conditioning=None, callback=None, normals_sequence=None,
               img_callback=None, quantize_x0=False, eta=0., mask=None, x0=None, temperature=1.,
               noise_dropout=0., score_corrector=None, corrector_kwargs=None, verbose=True, x_T=None,
               log_every_t=100, unconditional_guidance_scale=1., unconditional_conditioning=None, **kwargs):
        device = self.model.betas.device
        C_, H, W = shape
        img = (torch.randn((batch_size, C_, H, W), device=device) if x_T is None
               else x_T.to(device=device, dtype=torch.float32).clone())
        sch = dpm_solver_pp_2m_schedule(self.alphas_cumprod, S)
        unet = self._unet()
        cfg = not (unconditional_conditioning is None or unconditional_guidance_scale == 1.)
        if unet is not None and 2 * batch_size <= unet.max_batch and S <= 256 and callback is None and img_callback is None:
            cond = conditioning.to(device=device, dtype=torch.float32).contiguous()
            # without guidance the reference evaluates the conditional branch only (:333-334); the fused step
            # always carries two branches, so both get `cond` and the combine e_u + 1 * (e_c - e_u) returns e_c
            unc = (unconditional_conditioning if cfg else conditioning).to(device=device, dtype=torch.float32).contiguous()
            scale = float(unconditional_guidance_scale) if cfg else 1.0
            if (C_, H, W) != (unet.in_channels, *unet.latent_size) or tuple(img.shape) != (batch_size, C_, H, W):
                raise ValueError(f"x_T / shape do not match the UNet latent {(unet.in_channels, *unet.latent_size)}")
            if (cond.dim() != 3 or cond.shape != unc.shape or cond.shape[0] != batch_size or
                    cond.shape[1] > unet.max_context_len or cond.shape[2] != unet.context_dim):
                raise ValueError(f"conditioning must be [{batch_size}, L <= {unet.max_context_len}, {unet.context_dim}]")
            h = unet.engine(device)
            fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
            with torch.cuda.device(device):
                L.check(L.lib().dfb_dpm_solver_sample(
                    h, L.ptr(img), L.ptr(cond), L.ptr(unc), batch_size, cond.shape[1], scale, S, fp(sch["t_input"]),
                    fp(sch["sigma"]), fp(sch["alpha"]), fp(sch["cx"]), fp(sch["a_coef"]), fp(sch["inv_r0"]),
                    sch["order"].ctypes.data_as(C.POINTER(C.c_int32)), None, L.cur_stream()), "dfb_dpm_solver_sample")
            return img, None
        return self._host_loop(img, sch, S, conditioning, unconditional_guidance_scale, unconditional_conditioning,
                               None, None, 0.0, callback), None

    @torch.no_grad()
    def sample_with_classifier(self, S, batch_size, shape, conditioning=None, origin_cond=None, x_T=None,
                               unconditional_guidance_scale=1., unconditional_conditioning=None, classifier=None,
                               classifier_guide_scale=0.0, callback=None, **kwargs):
        device = self.model.betas.device
        C_, H, W = shape
        img = (torch.randn((batch_size, C_, H, W), device=device) if x_T is None
               else x_T.to(device=device, dtype=torch.float32).clone())
        sch = dpm_solver_pp_2m_schedule(self.alphas_cumprod, S)
        return self._host_loop(img, sch, S, conditioning, unconditional_guidance_scale, unconditional_conditioning,
                               origin_cond, classifier, classifier_guide_scale, callback), None

    def _host_loop(self, x, sch, S, cond, scale, uncond, origin_cond, classifier, cscale, callback):
        """The same multistep updates with `apply_model` per evaluation (and the classifier gradient of
        dpm_solver.py:1340-1390 when given); elementwise math in fp32 torch, the reference's operation order."""
        b, dev = x.shape[0], x.device
        f = lambda v: torch.tensor(float(v), dtype=torch.float32, device=dev)
        cfg = not (uncond is None or scale == 1.)
        m_prev = None
        for k in range(S):
            t_in = torch.full((b,), float(sch["t_input"][k]), dtype=torch.float32, device=dev)
            if cfg:
                e = self.model.apply_model(torch.cat([x] * 2), torch.cat([t_in] * 2), torch.cat([uncond, cond])).float()
                e_u, e_c = e.chunk(2)
                noise = e_u + scale * (e_c - e_u)
                if classifier is not None:
                    if hasattr(classifier, "loglikelihood_grad"):
                        grad = classifier.loglikelihood_grad(x, t_in, origin_cond, 1.0)
                    else:
                        with torch.enable_grad():
                            x_in = x.detach().requires_grad_(True)
                            grad = torch.autograd.grad(torch.log(classifier(x_in, t=t_in, video_feat=origin_cond)).sum(), x_in)[0]
                    noise = noise - cscale * f(sch["sigma"][k]) * grad
            else:
                noise = self.model.apply_model(x, t_in, cond).float()
            m = (x - f(sch["sigma"][k]) * noise) / f(sch["alpha"][k])
            xn = f(sch["cx"][k]) * x - f(sch["a_coef"][k]) * m
            if int(sch["order"][k]) == 2:
                xn = xn - 0.5 * f(sch["a_coef"][k]) * (f(sch["inv_r0"][k]) * (m - m_prev))
            x, m_prev = xn, m
            if callback:
                callback(k)
        return x
