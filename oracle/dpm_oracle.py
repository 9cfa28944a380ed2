"""Oracle (TEST INFRASTRUCTURE ONLY -- imported by tests/ and nothing else): CPU restatement of the
reference's DPM-Solver++ multistep sampler as the notebook uses it and of its PLMS sampler.

DPM-Solver: diff_foley/models/diffusion/dpm_solver/sampler.py:25-87 builds
    NoiseScheduleVP('discrete', alphas_cumprod)                       dpm_solver.py:98-108
    model_wrapper(..., guidance_type="classifier-free")              dpm_solver.py:278-344
    DPM_Solver(model_fn, ns, predict_x0=True).sample(x, steps=S, skip_type="time_uniform",
        method="multistep", order=2, lower_order_final=True)          dpm_solver.py:1064-1096
PLMS: diff_foley/models/diffusion/plms.py:172-236 (p_sample_plms) on the DDIM schedule of :24-55.

Pinned by tests/test_oracle.py against tests/golden/dpm_small.npz / plms_small.npz, which
tests/golden/make_golden.py produced by running the reference's own DPMSolverSampler / PLMSSampler.
Everything is fp32 torch on CPU, plain loops, same operation order as the reference.
"""
import math

import torch

from . import ddim_oracle


# ------------------------------------------------------------------------------ noise schedule
def interpolate(x, xp, yp):
    """Piecewise-linear y(x) through the keypoints (xp, yp) (xp ascending), linear extrapolation from the
    outermost segment beyond the ends -- the function dpm_solver.py:1132-1171 evaluates with a sort."""
    K = xp.shape[0]
    out = torch.empty_like(x)
    for i in range(x.shape[0]):
        xi = x[i]
        # number of keypoints strictly below x (the reference's x_idx); segment = [j, j+1]
        n_below = int((xp < xi).sum())
        j = min(max(n_below - 1, 0), K - 2)
        out[i] = yp[j] + (xi - xp[j]) * (yp[j + 1] - yp[j]) / (xp[j + 1] - xp[j])
    return out


class NoiseScheduleDiscrete:
    """NoiseScheduleVP('discrete', alphas_cumprod=...) (dpm_solver.py:98-108, 125-156)."""

    def __init__(self, alphas_cumprod):
        self.log_alpha = 0.5 * torch.log(alphas_cumprod.to(torch.float32))
        self.total_N = self.log_alpha.shape[0]
        self.T = 1.0
        self.t_array = torch.linspace(0., 1., self.total_N + 1)[1:]

    def marginal_log_mean_coeff(self, t):
        return interpolate(t, self.t_array, self.log_alpha)

    def marginal_alpha(self, t):
        return torch.exp(self.marginal_log_mean_coeff(t))

    def marginal_std(self, t):
        return torch.sqrt(1. - torch.exp(2. * self.marginal_log_mean_coeff(t)))

    def marginal_lambda(self, t):
        lm = self.marginal_log_mean_coeff(t)
        return lm - 0.5 * torch.log(1. - torch.exp(2. * lm))

    def model_input_time(self, t):
        """dpm_solver.py:278-287: continuous t in [1/N, 1] -> the fractional fp32 `timesteps` the UNet gets."""
        return (t - 1. / self.total_N) * 1000.


# ---------------------------------------------------------------------------------- DPM-Solver++
@torch.no_grad()
def dpm_solver_sample(eps_fn, x_T, cond, uncond, scale, steps, alphas_cumprod=None, order=2,
                      lower_order_final=True):
    """x_T -> x_0.  eps_fn(x_in [2B,...], t_in fp32 [2B], c_in [2B,L,D]) -> eps [2B,...] (apply_model)."""
    ns = NoiseScheduleDiscrete(ddim_oracle.alphas_cumprod() if alphas_cumprod is None else alphas_cumprod)
    b = x_T.shape[0]
    t_0, t_T = 1. / ns.total_N, ns.T
    timesteps = torch.linspace(t_T, t_0, steps + 1)              # 'time_uniform', dpm_solver.py:427-428

    def model_fn(x, vec_t):                                        # data prediction, :386-394 + :321-344
        t_in = ns.model_input_time(vec_t)
        out = eps_fn(torch.cat([x, x]), torch.cat([t_in, t_in]), torch.cat([uncond, cond]))
        e_u, e_c = out.chunk(2)
        noise = e_u + scale * (e_c - e_u)
        alpha_t, sigma_t = ns.marginal_alpha(vec_t), ns.marginal_std(vec_t)
        ex = lambda v: v[:, None, None, None]
        return (x - ex(sigma_t) * noise) / ex(alpha_t)

    def first_update(x, s, t, model_s):                            # :504-533 (predict_x0 branch)
        h = ns.marginal_lambda(t) - ns.marginal_lambda(s)
        sigma_s, sigma_t = ns.marginal_std(s), ns.marginal_std(t)
        alpha_t = torch.exp(ns.marginal_log_mean_coeff(t))
        phi_1 = torch.expm1(-h)
        ex = lambda v: v[:, None, None, None]
        return ex(sigma_t / sigma_s) * x - ex(alpha_t * phi_1) * model_s

    def second_update(x, models, ts, t):                           # :755-790 ('dpm_solver' type, predict_x0)
        m1, m0 = models
        t1, t0 = ts
        l1, l0, lt = ns.marginal_lambda(t1), ns.marginal_lambda(t0), ns.marginal_lambda(t)
        sigma_0, sigma_t = ns.marginal_std(t0), ns.marginal_std(t)
        alpha_t = torch.exp(ns.marginal_log_mean_coeff(t))
        h_0, h = l0 - l1, lt - l0
        r0 = h_0 / h
        ex = lambda v: v[:, None, None, None]
        D1 = ex(1. / r0) * (m0 - m1)
        return (ex(sigma_t / sigma_0) * x - ex(alpha_t * (torch.exp(-h) - 1.)) * m0
                - 0.5 * ex(alpha_t * (torch.exp(-h) - 1.)) * D1)

    x = x_T.clone()
    vec = lambda i: timesteps[i].expand(b)
    models, ts = [model_fn(x, vec(0))], [vec(0)]                   # :1069-1071
    for init_order in range(1, order):                             # :1073-1077
        x = first_update(x, ts[-1], vec(init_order), models[-1])
        models.append(model_fn(x, vec(init_order)))
        ts.append(vec(init_order))
    for step in range(order, steps + 1):                           # :1079-1092
        step_order = min(order, steps + 1 - step) if (lower_order_final and steps < 15) else order
        if step_order == 1:
            x = first_update(x, ts[-1], vec(step), models[-1])
        else:
            x = second_update(x, models, ts, vec(step))
        ts[0], models[0] = ts[1], models[1]
        ts[-1] = vec(step)
        if step < steps:
            models[-1] = model_fn(x, vec(step))
    return x


# ----------------------------------------------------------------------------------------- PLMS
@torch.no_grad()
def plms_sample(eps_fn, x_T, cond, uncond, scale, num_steps):
    """PLMSSampler.plms_sampling / p_sample_plms (plms.py:118-236), eta = 0, on the DDIM sub-schedule."""
    c = ddim_oracle.ddim_coefficients(num_steps)
    x = x_T.clone()
    b = x.shape[0]
    steps = c["timesteps"]
    old_eps = []

    def model_out(xx, t_int):
        ts = torch.full((b,), int(t_int), dtype=torch.long)
        out = eps_fn(torch.cat([xx, xx]), torch.cat([ts, ts]), torch.cat([uncond, cond]))
        e_u, e_c = out.chunk(2)
        return e_u + scale * (e_c - e_u)

    pred = None
    for i, step in enumerate(steps):
        t_next = steps[min(i + 1, len(steps) - 1)]                 # plms.py:139-140
        e_t = model_out(x, step)
        if len(old_eps) == 0:
            x_prev, _ = ddim_oracle.ddim_step(x, None, e_t, 1.0, c, i)
            e_next = model_out(x_prev, t_next)
            e_prime = (e_t + e_next) / 2
        elif len(old_eps) == 1:
            e_prime = (3 * e_t - old_eps[-1]) / 2
        elif len(old_eps) == 2:
            e_prime = (23 * e_t - 16 * old_eps[-1] + 5 * old_eps[-2]) / 12
        else:
            e_prime = (55 * e_t - 59 * old_eps[-1] + 37 * old_eps[-2] - 9 * old_eps[-3]) / 24
        x, pred = ddim_oracle.ddim_step(x, None, e_prime, 1.0, c, i)
        old_eps.append(e_t)
        if len(old_eps) >= 4:
            old_eps.pop(0)
    return x, pred
