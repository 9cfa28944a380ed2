"""Generates tests/golden/*.npz by running the REFERENCE's own modules (imported unmodified from
/root/reference through the import shims in tests/golden/_shims) on seeded inputs and seeded
weights (oracle.unet_oracle.seeded_state_dict).  Run in the build container only:

    python tests/golden/make_golden.py            # all fixtures  (~4 min on 8 vCPU)
    python tests/golden/make_golden.py small      # just the reduced-width ones
    python tests/golden/make_golden.py r2         # the round-2 additions (cond embed, DPM-Solver, PLMS, S=30,
                                                  # full-width classifier-guided DDIM)

The reference ships no tests or golden vectors of its own, so these files are what pins
the oracle (oracle/) and, through it and directly, the CUDA path.  Nothing here is read at run time
on the GPU box except the .npz outputs.
"""
import os
import sys
import time

import numpy as np
import torch
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("DIFF_FOLEY_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "_shims"))
sys.path.insert(0, REF)

from oracle import ddim_oracle, unet_oracle  # noqa: E402

torch.manual_seed(0)


def ref_unet(cfg):
    from diff_foley.modules.diffusionmodules.openai_unetmodel import UNetModel
    m = UNetModel(image_size=32, in_channels=cfg["in_channels"], out_channels=cfg["out_channels"],
                  model_channels=cfg["model_channels"],
                  attention_resolutions=list(cfg["attention_resolutions"]),
                  num_res_blocks=cfg["num_res_blocks"], channel_mult=list(cfg["channel_mult"]),
                  num_heads=cfg["num_heads"], use_spatial_transformer=True, transformer_depth=1,
                  context_dim=cfg["context_dim"], use_checkpoint=True, legacy=False)
    return m.eval()


def unet_inputs(cfg, b_eff, seed, t_values):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(b_eff, cfg["in_channels"], cfg["latent_h"], cfg["latent_w"], generator=g)
    ctx = torch.randn(b_eff, cfg["context_len"], cfg["context_dim"], generator=g)
    ctx[: b_eff // 2] = 0  # the uncond half of a CFG batch is all-zero context (notebook cell 13)
    t = torch.tensor(t_values, dtype=torch.long)
    return x, t, ctx


def gen_unet(name, cfg, seed, b_eff, t_values, taps=()):
    t0 = time.time()
    sd = unet_oracle.seeded_state_dict(cfg, seed)
    m = ref_unet(cfg)
    missing, unexpected = m.load_state_dict(sd, strict=True), None
    x, t, ctx = unet_inputs(cfg, b_eff, seed + 1000, t_values)
    acts = {}
    hooks = []
    for tap in taps:
        mod = m.get_submodule(tap)
        hooks.append(mod.register_forward_hook(lambda _m, _i, o, tap=tap: acts.__setitem__(tap, o.detach().numpy())))
    with torch.no_grad():
        eps = m(x, t, context=ctx)
    for h in hooks:
        h.remove()
    out = dict(x=x.numpy(), t=t.numpy(), ctx=ctx.numpy(), eps=eps.numpy(), seed=np.int64(seed))
    for k, v in acts.items():
        out["tap:" + k] = v
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: eps absmax {float(eps.abs().max()):.4f} rms {float(eps.pow(2).mean().sqrt()):.4f} "
          f"({time.time() - t0:.1f}s)")


class _StubLDM:
    """The attributes DDIMSampler touches on the model (ddim.py:15-56, 231-273), around the
    reference UNet; apply_model has the semantics of ddpm.py:925-1026 for the crossattn key."""

    def __init__(self, unet):
        self.unet = unet
        ac = ddim_oracle.alphas_cumprod()
        betas = torch.linspace(ddim_oracle.LINEAR_START ** 0.5, ddim_oracle.LINEAR_END ** 0.5,
                               ddim_oracle.NUM_TIMESTEPS, dtype=torch.float64) ** 2
        self.num_timesteps = ddim_oracle.NUM_TIMESTEPS
        self.betas = betas.float()
        self.alphas_cumprod = ac
        self.alphas_cumprod_prev = torch.cat([torch.ones(1), ac[:-1]])
        self.device = torch.device("cpu")
        self.parameterization = "eps"

    def apply_model(self, x, t, c):
        return self.unet(x, t, context=c)


def gen_ddim(name, cfg, seed, n_clips, steps, scale, use_real_ldm=False):
    """Runs the reference's DDIMSampler.sample (ddim.py:58-113) on CPU."""
    t0 = time.time()
    from diff_foley.models.diffusion.ddim import DDIMSampler

    class CpuDDIM(DDIMSampler):  # ddim.py:21-25 hard-codes .to("cuda")
        def register_buffer(self, name, attr):
            setattr(self, name, attr)

    sd = unet_oracle.seeded_state_dict(cfg, seed)
    unet = ref_unet(cfg)
    unet.load_state_dict(sd, strict=True)
    if use_real_ldm:
        # the real LatentDiffusion wrapper (ddpm.py:434) around the same UNet: proves the stub's
        # schedule/apply_model equal the reference's
        with open(os.path.join(REF, "inference/config/Stage2_LDM.yaml")) as f:
            y = yaml.safe_load(f)["model"]
        from diff_foley.util import instantiate_from_config
        y["params"]["unet_config"]["params"].update(
            model_channels=cfg["model_channels"], channel_mult=list(cfg["channel_mult"]),
            num_heads=cfg["num_heads"], context_dim=cfg["context_dim"],
            attention_resolutions=list(cfg["attention_resolutions"]))
        ldm = instantiate_from_config(y).eval()
        ldm.model.diffusion_model.load_state_dict(sd, strict=True)
        model = ldm
    else:
        model = _StubLDM(unet)
    g = torch.Generator().manual_seed(seed + 2000)
    x_T = torch.randn(n_clips, cfg["in_channels"], cfg["latent_h"], cfg["latent_w"], generator=g)
    cond = torch.randn(n_clips, cfg["context_len"], cfg["context_dim"], generator=g)
    uncond = torch.zeros_like(cond)
    sampler = CpuDDIM(model)
    with torch.no_grad():
        samples, inter = sampler.sample(S=steps, batch_size=n_clips,
                                        shape=(cfg["in_channels"], cfg["latent_h"], cfg["latent_w"]),
                                        conditioning=cond, eta=0.0, verbose=False, x_T=x_T,
                                        unconditional_guidance_scale=scale,
                                        unconditional_conditioning=uncond)
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"), x_T=x_T.numpy(), cond=cond.numpy(), samples=samples.numpy(),
        pred_x0=inter["pred_x0"][-1].numpy(), seed=np.int64(seed), steps=np.int64(steps),
        scale=np.float32(scale), ddim_timesteps=np.asarray(sampler.ddim_timesteps),
        ddim_alphas=np.asarray(sampler.ddim_alphas), ddim_alphas_prev=np.asarray(sampler.ddim_alphas_prev),
        ddim_sqrt_one_minus_alphas=np.asarray(sampler.ddim_sqrt_one_minus_alphas),
        ddim_sigmas=np.asarray(sampler.ddim_sigmas))
    print(f"{name}: latent rms {float(samples.pow(2).mean().sqrt()):.4f} ({time.time() - t0:.1f}s)")


def ref_classifier(cfg):
    from diff_foley.modules.double_guidance.alignment_backbone import Classifier_Backbone
    m = Classifier_Backbone(image_size=32, in_channels=cfg["in_channels"], out_channels=cfg["out_channels"],
                            model_channels=cfg["model_channels"],
                            attention_resolutions=list(cfg["attention_resolutions"]),
                            num_res_blocks=cfg["num_res_blocks"], channel_mult=list(cfg["channel_mult"]),
                            num_heads=cfg["num_heads"], use_spatial_transformer=True, transformer_depth=1,
                            context_dim=cfg["context_dim"], use_checkpoint=True, legacy=False)
    return m.eval()


def gen_classifier(name, cfg, seed, b, ctx_len):
    """Classifier_Backbone probabilities + the reference's own cal_classifier_loglikelihood_grad."""
    from diff_foley.models.diffusion.ddim import DDIMSampler
    from oracle import classifier_oracle
    sd = classifier_oracle.seeded_state_dict(cfg, seed)
    m = ref_classifier(cfg)
    m.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(seed + 50)
    x = torch.randn(b, 4, 16, 64, generator=g)
    feats = torch.nn.functional.normalize(torch.randn(b, ctx_len, cfg["context_dim"], generator=g), dim=-1)
    t = torch.tensor([961, 41, 500, 1][:b], dtype=torch.long)
    with torch.no_grad():
        prob = m(x, timesteps=t, context=feats)
    wrapper = lambda x_in, t, video_feat: m(x_in, timesteps=t, context=video_feat)
    grad = DDIMSampler.cal_classifier_loglikelihood_grad(None, wrapper, x, t, feats, classifier_guide_scale=50.0)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), x=x.numpy(), t=t.numpy(), feats=feats.numpy(),
                        prob=prob.numpy(), grad=grad.numpy(), seed=np.int64(seed))
    print(f"{name}: prob {prob.flatten().tolist()} grad rms {float(grad.pow(2).mean().sqrt()):.4e}")


def gen_ddim_classifier(name, ucfg, ccfg, seed, n_clips, steps, scale, cscale):
    """DDIMSampler.sample_with_classifier (ddim.py:115-176) on CPU: CFG + classifier guidance."""
    t0 = time.time()
    from diff_foley.models.diffusion.ddim import DDIMSampler
    from oracle import classifier_oracle

    class CpuDDIM(DDIMSampler):
        def register_buffer(self, name, attr):
            setattr(self, name, attr)

    unet = ref_unet(ucfg)
    unet.load_state_dict(unet_oracle.seeded_state_dict(ucfg, seed), strict=True)
    clf = ref_classifier(ccfg)
    clf.load_state_dict(classifier_oracle.seeded_state_dict(ccfg, seed + 1), strict=True)
    classifier = lambda x_in, t, video_feat: clf(x_in, timesteps=t, context=video_feat)
    g = torch.Generator().manual_seed(seed + 3000)
    x_T = torch.randn(n_clips, 4, ucfg["latent_h"], ucfg["latent_w"], generator=g)
    cond = torch.randn(n_clips, ucfg["context_len"], ucfg["context_dim"], generator=g)
    feats = torch.nn.functional.normalize(torch.randn(n_clips, 33, ccfg["context_dim"], generator=g), dim=-1)
    sampler = CpuDDIM(_StubLDM(unet))
    samples, _ = sampler.sample_with_classifier(
        S=steps, batch_size=n_clips, shape=(4, ucfg["latent_h"], ucfg["latent_w"]), conditioning=cond,
        origin_cond=feats, eta=0.0, verbose=False, x_T=x_T, unconditional_guidance_scale=scale,
        unconditional_conditioning=torch.zeros_like(cond), classifier=classifier, classifier_guide_scale=cscale)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), x_T=x_T.numpy(), cond=cond.numpy(), feats=feats.numpy(),
                        samples=samples.numpy(), seed=np.int64(seed), steps=np.int64(steps),
                        scale=np.float32(scale), cscale=np.float32(cscale))
    print(f"{name}: latent rms {float(samples.pow(2).mean().sqrt()):.4f} ({time.time() - t0:.1f}s)")


def gen_sampler(name, kind, cfg, seed, n_clips, steps, scale):
    """The reference's DPMSolverSampler.sample (dpm_solver/sampler.py:25-87: DPM-Solver++ multistep order 2,
    the notebook's default) or PLMSSampler.sample (plms.py:57-112) on CPU, CFG `scale`."""
    t0 = time.time()
    if kind == "dpm":
        from diff_foley.models.diffusion.dpm_solver import DPMSolverSampler as Ref
    else:
        from diff_foley.models.diffusion.plms import PLMSSampler as Ref

    class Cpu(Ref):  # register_buffer hard-codes .to("cuda") (sampler.py:18-22, plms.py:17-21)
        def register_buffer(self, name, attr):
            setattr(self, name, attr)

    unet = ref_unet(cfg)
    unet.load_state_dict(unet_oracle.seeded_state_dict(cfg, seed), strict=True)
    g = torch.Generator().manual_seed(seed + 4000)
    x_T = torch.randn(n_clips, cfg["in_channels"], cfg["latent_h"], cfg["latent_w"], generator=g)
    cond = torch.randn(n_clips, cfg["context_len"], cfg["context_dim"], generator=g)
    sampler = Cpu(_StubLDM(unet))
    with torch.no_grad():
        samples, _ = sampler.sample(S=steps, batch_size=n_clips,
                                    shape=(cfg["in_channels"], cfg["latent_h"], cfg["latent_w"]), conditioning=cond,
                                    eta=0.0, verbose=False, x_T=x_T.clone(), unconditional_guidance_scale=scale,
                                    unconditional_conditioning=torch.zeros_like(cond))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), x_T=x_T.numpy(), cond=cond.numpy(), samples=samples.numpy(),
                        seed=np.int64(seed), steps=np.int64(steps), scale=np.float32(scale))
    print(f"{name}: latent rms {float(samples.pow(2).mean().sqrt()):.4f} ({time.time() - t0:.1f}s)")


def gen_cond_embed(name="cond_embed", seed=41):
    """Video_Feat_Encoder_Posembed.forward (cond_stage/video_feat_encoder.py:4-18) with the Stage2_LDM.yaml
    parameters (origin_dim 512, embed_dim 768, seq_len 40) on seeded weights and unit-norm CAVP-like features."""
    from diff_foley.modules.cond_stage.video_feat_encoder import Video_Feat_Encoder_Posembed
    m = Video_Feat_Encoder_Posembed(origin_dim=512, embed_dim=768, seq_len=40).eval()
    sd = ddim_oracle.cond_stage_seeded_state(seed)
    m.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(seed + 1)
    feats = torch.nn.functional.normalize(torch.randn(3, 32, 512, generator=g), dim=-1)
    with torch.no_grad():
        out = m(feats)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), feats=feats.numpy(), out=out.numpy(), seed=np.int64(seed))
    print(f"{name}: out {tuple(out.shape)} rms {float(out.pow(2).mean().sqrt()):.4f}")


SMALL = unet_oracle.small_unet_cfg()                       # 64 ch, heads 4 -> head dims 16/32/64
SMALL_ODD = unet_oracle.small_unet_cfg(model_channels=128, channel_mult=(1, 2), num_heads=8,
                                       context_dim=64, latent_h=8, latent_w=16, context_len=33,
                                       attention_resolutions=(2, 1))  # head dims 16/32, ragged ctx
FULL = unet_oracle.DIFF_FOLEY_UNET

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    torch.set_num_threads(os.cpu_count())
    taps = ("input_blocks.1", "input_blocks.3", "input_blocks.4", "middle_block", "output_blocks.2",
            "output_blocks.5", "output_blocks.11")
    from oracle import classifier_oracle
    CLF_SMALL = dict(classifier_oracle.DIFF_FOLEY_CLASSIFIER, model_channels=64, num_heads=4, context_dim=64)
    if which in ("all", "small"):
        gen_unet("unet_small", SMALL, 1, 2, [961, 961], taps)
        gen_unet("unet_small_b3", SMALL, 2, 3, [41, 500, 999])
        gen_unet("unet_small_odd", SMALL_ODD, 3, 2, [1, 1], ("input_blocks.1", "middle_block"))
        gen_ddim("ddim_small", SMALL, 1, 2, 25, 4.5)
        gen_ddim("ddim_small_ldm", SMALL, 4, 1, 5, 4.5, use_real_ldm=True)
        gen_classifier("classifier_small", CLF_SMALL, 5, 3, 33)
        gen_classifier("classifier_full", classifier_oracle.DIFF_FOLEY_CLASSIFIER, 6, 2, 33)
        gen_ddim_classifier("ddim_classifier_small", SMALL, CLF_SMALL, 9, 2, 25, 4.5, 50.0)
    if which in ("all", "r2"):
        gen_cond_embed()
        gen_ddim("ddim_small_s30", SMALL, 11, 1, 30, 4.5)      # 1000 % 30 != 0: 34 schedule entries, all run
        gen_sampler("dpm_small", "dpm", SMALL, 12, 2, 25, 4.5)
        gen_sampler("dpm_small_s10", "dpm", SMALL, 13, 1, 10, 4.5)   # steps < 15: lower-order final step
        gen_sampler("plms_small", "plms", SMALL, 14, 2, 25, 4.5)
        # config 3 at FULL width: CFG 4.5 + classifier guidance 50, 5 steps (1, 201, ..., 801)
        gen_ddim_classifier("ddim_classifier_full", FULL, classifier_oracle.DIFF_FOLEY_CLASSIFIER, 15, 1, 5, 4.5, 50.0)
    if which == "all":
        gen_unet("unet_full", FULL, 7, 2, [961, 961])
        gen_unet("unet_full_t41", FULL, 7, 2, [41, 41])
        gen_ddim("ddim_full", FULL, 7, 1, 25, 4.5)


def gen_vae_decode(name="vae_decode", seed=21):
    """decoded image of the reference-sampled latent (ddim_full.npz) through the reference's own
    AutoencoderKL.decode with seeded decoder weights: the north-star tolerance is stated on mel =
    channel 0 of this."""
    from diff_foley.models.autoencoder import AutoencoderKL
    from oracle import vae_oracle
    with open(os.path.join(REF, "inference/config/Stage2_LDM.yaml")) as f:
        y = yaml.safe_load(f)["model"]["params"]["first_stage_config"]["params"]
    vae = AutoencoderKL(**y).eval()
    sd = vae_oracle.seeded_state_dict(seed)
    missing, unexpected = vae.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.startswith(("encoder.", "quant_conv.", "loss.")) for k in missing), (missing, unexpected)
    z = torch.from_numpy(np.load(os.path.join(HERE, "ddim_full.npz"))["samples"])
    with torch.no_grad():
        img = vae.decode(z / 0.18215)             # decode_first_stage, ddpm.py:739-797
    np.savez_compressed(os.path.join(HERE, name + ".npz"), z=z.numpy(), mel=img[:, 0].numpy(),
                        img_rms=np.float32(img.pow(2).mean().sqrt()), seed=np.int64(seed))
    print(f"{name}: image {tuple(img.shape)} mel rms {float(img[:, 0].pow(2).mean().sqrt()):.4f}")


if __name__ == "__main__" and (len(sys.argv) < 2 or sys.argv[1] in ("all", "vae")):
    gen_vae_decode()


def gen_cavp(name, seed, B, T, HW, spec_T):
    """CAVP_Inference.encode_video / encode_spec (inference/model/cavp_model.py:47-84) through the mmcv
    import shim, seeded weights incl. randomised BatchNorm statistics.  Inputs are regenerated from the
    seed by the tests (a 224x224x32 clip is 19 MB), only the outputs are stored."""
    t0 = time.time()
    sys.path.insert(0, os.path.join(REF, "inference"))
    from model.cavp_model import CAVP_Inference
    from oracle import cavp_oracle
    m = CAVP_Inference("Slowonly_pool", "cnn14_pool", 512).eval()
    sd = cavp_oracle.seeded_state_dict(seed)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all("num_batches_tracked" in k or k == "logit_scale" for k in missing), (missing, unexpected)
    g = torch.Generator().manual_seed(seed + 77)
    video = torch.rand(B, T, 3, HW, HW, generator=g)
    spec = torch.randn(B, 128, spec_T, generator=g)
    with torch.no_grad():
        v = m.encode_video(video, normalize=True, pool=False)
        v_raw = m.encode_video(video, normalize=False, pool=False)
        s = m.encode_spec(spec, normalize=True, pool=False)
        s_raw = m.encode_spec(spec, normalize=False, pool=False)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), video_feat=v.numpy(), video_raw=v_raw.numpy(),
                        spec_feat=s.numpy(), spec_raw=s_raw.numpy(), seed=np.int64(seed),
                        shape=np.asarray([B, T, HW, spec_T]))
    print(f"{name}: video {tuple(v.shape)} raw rms {float(v_raw.pow(2).mean().sqrt()):.3f}  spec {tuple(s.shape)} "
          f"raw rms {float(s_raw.pow(2).mean().sqrt()):.3f} ({time.time() - t0:.1f}s)")


if __name__ == "__main__" and (len(sys.argv) < 2 or sys.argv[1] in ("all", "cavp")):
    gen_cavp("cavp_small", 31, 2, 4, 64, 64)
    gen_cavp("cavp_full", 32, 1, 32, 224, 512)
