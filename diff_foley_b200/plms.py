"""PLMSSamplerB200 -- the reference's PLMSSampler interface (diff_foley/models/diffusion/plms.py:12-236) on
the CUDA engine: pseudo linear multistep (Adams-Bashforth on eps, plms.py:216-234) over the DDIM
sub-schedule, eta = 0.  The UNet evaluations go through `model.apply_model` (UNetModelB200 -> libdfb.so), the
x_prev / pred_x0 update is the fused `dfb_ddim_step` kernel; the eps history combination is four tiny
elementwise ops per step on the host loop (the first step needs a second model evaluation, :217-220, so
the step is not a fixed graph).
"""
import torch

from . import _lib as L
from .ddim import DDIMSamplerB200


class PLMSSamplerB200(DDIMSamplerB200):
    def make_schedule(self, ddim_num_steps, ddim_discretize="uniform", ddim_eta=0., verbose=False):
        if ddim_eta != 0:
            raise ValueError("ddim_eta must be 0 for PLMS")     # plms.py:24-25
        return super().make_schedule(ddim_num_steps, ddim_discretize, ddim_eta, verbose)

    @torch.no_grad()
    def sample(self, S, batch_size, shape, conditioning=None, callback=None, normals_sequence=None,
               img_callback=None, quantize_x0=False, eta=0., mask=None, x0=None, temperature=1.,
               noise_dropout=0., score_corrector=None, corrector_kwargs=None, verbose=True, x_T=None,
               log_every_t=100, unconditional_guidance_scale=1., unconditional_conditioning=None, **kwargs):
        if quantize_x0 or mask is not None or x0 is not None or score_corrector is not None:
            raise NotImplementedError("quantize_x0 / mask / score_corrector are outside the hot path")
        self.make_schedule(ddim_num_steps=S, ddim_eta=eta, verbose=verbose)
        st = self._steps
        device = self.model.betas.device
        C_, H, W = shape
        img = (torch.randn((batch_size, C_, H, W), device=device) if x_T is None
               else x_T.to(device=device, dtype=torch.float32).clone())
        intermediates = {"x_inter": [img], "pred_x0": [img]}
        cfg = not (unconditional_conditioning is None or unconditional_guidance_scale == 1.)
        steps = st["timesteps"]
        total = len(steps)
        lib = L.lib()

        def model_out(x, t_int):
            ts = torch.full((batch_size,), int(t_int), device=device, dtype=torch.long)
            if not cfg:
                return self.model.apply_model(x, ts, conditioning).float()
            e = self.model.apply_model(torch.cat([x] * 2), torch.cat([ts] * 2),
                                       torch.cat([unconditional_conditioning, conditioning])).float()
            e_u, e_c = e.chunk(2)
            return e_u + unconditional_guidance_scale * (e_c - e_u)

        def update(x, e, i):
            nxt, pred = torch.empty_like(x), torch.empty_like(x)
            e = e.contiguous()
            with torch.cuda.device(device):
                L.check(lib.dfb_ddim_step(L.ptr(x), None, L.ptr(e), None, 1.0, float(st["sqrt_one_minus_at"][i]),
                                          float(st["sqrt_at"][i]), float(st["sqrt_a_prev"][i]), float(st["dir_coef"][i]),
                                          0.0, L.ptr(nxt), L.ptr(pred), x.numel(), L.cur_stream()), "dfb_ddim_step")
            return nxt, pred

        old_eps = []
        for i, step in enumerate(steps):
            index = total - i - 1
            t_next = steps[min(i + 1, total - 1)]
            e_t = model_out(img, step)
            if len(old_eps) == 0:        # pseudo improved Euler (plms.py:216-220)
                x_prev, _ = update(img, e_t, i)
                e_prime = (e_t + model_out(x_prev, t_next)) / 2
            elif len(old_eps) == 1:
                e_prime = (3 * e_t - old_eps[-1]) / 2
            elif len(old_eps) == 2:
                e_prime = (23 * e_t - 16 * old_eps[-1] + 5 * old_eps[-2]) / 12
            else:
                e_prime = (55 * e_t - 59 * old_eps[-1] + 37 * old_eps[-2] - 9 * old_eps[-3]) / 24
            img, pred = update(img, e_prime, i)
            old_eps.append(e_t)
            if len(old_eps) >= 4:
                old_eps.pop(0)
            if callback: callback(i)
            if img_callback: img_callback(pred, i)
            if index % log_every_t == 0 or index == total - 1:
                intermediates["x_inter"].append(img)
                intermediates["pred_x0"].append(pred)
        return img, intermediates
