"""CPU tests: the oracle (oracle/) against the golden vectors produced by the reference's own
modules (tests/golden/make_golden.py).  These pin the oracle; the GPU tests then compare the CUDA
path with the oracle and with the same golden files."""
import os

import numpy as np
import pytest
import torch

from oracle import ddim_oracle, unet_oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SMALL = unet_oracle.small_unet_cfg()
SMALL_ODD = unet_oracle.small_unet_cfg(model_channels=128, channel_mult=(1, 2), num_heads=8,
                                       context_dim=64, latent_h=8, latent_w=16, context_len=33,
                                       attention_resolutions=(2, 1))
FULL = unet_oracle.DIFF_FOLEY_UNET


def load(name):
    path = os.path.join(GOLD, name + ".npz")
    if not os.path.exists(path):
        pytest.skip(f"{name}.npz not generated")
    return np.load(path)


def rel_l2(a, b):
    a, b = torch.as_tensor(a).double().flatten(), torch.as_tensor(b).double().flatten()
    return float((a - b).norm() / b.norm())


@pytest.mark.parametrize("name,cfg", [("unet_small", SMALL), ("unet_small_b3", SMALL),
                                      ("unet_small_odd", SMALL_ODD)])
def test_unet_oracle_matches_reference(name, cfg):
    g = load(name)
    sd = unet_oracle.seeded_state_dict(cfg, int(g["seed"]))
    taps = {}
    eps = unet_oracle.unet_forward(sd, cfg, torch.from_numpy(g["x"]), torch.from_numpy(g["t"]),
                                   torch.from_numpy(g["ctx"]), taps)
    # same fp32 ATen kernels in a slightly different call order: agreement to a few ulp
    assert rel_l2(eps, g["eps"]) < 2e-6
    for k in g.files:
        if k.startswith("tap:"):
            assert rel_l2(taps[k[4:]], g[k]) < 2e-6, k


def test_unet_oracle_full_size():
    """The real Diff-Foley configuration (859.5 M parameters), one CFG step at t = 961."""
    g = load("unet_full")
    sd = unet_oracle.seeded_state_dict(FULL, int(g["seed"]))
    assert sum(v.numel() for v in sd.values()) == 859_520_964
    eps = unet_oracle.unet_forward(sd, FULL, torch.from_numpy(g["x"]), torch.from_numpy(g["t"]),
                                   torch.from_numpy(g["ctx"]))
    assert rel_l2(eps, g["eps"]) < 5e-6


def test_param_shapes_cover_reference_keys():
    shapes = unet_oracle.unet_param_shapes(FULL)
    assert len(shapes) == 686
    assert shapes["input_blocks.0.0.weight"] == (320, 4, 3, 3)
    assert shapes["output_blocks.0.0.skip_connection.weight"] == (1280, 2560, 1, 1)
    assert shapes["middle_block.1.transformer_blocks.0.attn2.to_k.weight"] == (1280, 768)
    assert shapes["output_blocks.5.2.conv.weight"] == (1280, 1280, 3, 3)
    assert shapes["output_blocks.2.1.conv.weight"] == (1280, 1280, 3, 3)


def test_schedule_matches_reference_sampler():
    g = load("ddim_small")
    c = ddim_oracle.ddim_coefficients(int(g["steps"]))
    ts = np.asarray(g["ddim_timesteps"])
    assert ts[0] == 1 and ts[-1] == 961 and len(ts) == 25
    assert np.array_equal(c["timesteps"], ts[::-1])
    # bit-exact fp32 scalars, in the reference's own order (ascending t) -> flip ours
    assert np.array_equal(c["a_t"][::-1], g["ddim_alphas"].astype(np.float32))
    assert np.array_equal(c["sqrt_one_minus_at"][::-1], g["ddim_sqrt_one_minus_alphas"].astype(np.float32))
    # the reference takes torch.full(...fp32).sqrt() of these (ddim.py:272); torch's CPU sqrt, not numpy's
    a_prev = torch.tensor(g["ddim_alphas_prev"], dtype=torch.float32)
    assert np.array_equal(c["sqrt_a_prev"][::-1], a_prev.sqrt().numpy())
    assert np.all(np.asarray(g["ddim_sigmas"]) == 0)


@pytest.mark.parametrize("name", ["ddim_small", "ddim_small_ldm"])
def test_ddim_oracle_matches_reference_sampler(name):
    g = load(name)
    sd = unet_oracle.seeded_state_dict(SMALL, int(g["seed"]))
    cond = torch.from_numpy(g["cond"])
    fn = lambda x, t, c: unet_oracle.unet_forward(sd, SMALL, x, t, c)
    x, pred = ddim_oracle.ddim_sample(fn, torch.from_numpy(g["x_T"]), cond, torch.zeros_like(cond),
                                      float(g["scale"]), int(g["steps"]))
    assert rel_l2(x, g["samples"]) < 2e-5
    assert rel_l2(pred, g["pred_x0"]) < 2e-5


def test_ddim_step_linearity():
    """size-independent property: the update is affine in (x, e) with the schedule's coefficients."""
    c = ddim_oracle.ddim_coefficients(25)
    g = torch.Generator().manual_seed(0)
    x, eu, ec = (torch.randn(3, 4, 16, 64, generator=g) for _ in range(3))
    xp, p0 = ddim_oracle.ddim_step(x, eu, ec, 4.5, c, 3)
    e = eu + 4.5 * (ec - eu)
    assert torch.allclose(p0 * float(c["sqrt_at"][3]) + float(c["sqrt_one_minus_at"][3]) * e, x, atol=1e-5)
    assert torch.allclose(xp, float(c["sqrt_a_prev"][3]) * p0 + float(c["dir_coef"][3]) * e, atol=1e-6)


# ----------------------------------------------------------------------------- classifier (row a15)
from oracle import classifier_oracle  # noqa: E402

CLF_SMALL = dict(classifier_oracle.DIFF_FOLEY_CLASSIFIER, model_channels=64, num_heads=4, context_dim=64)


@pytest.mark.parametrize("name,cfg", [("classifier_small", CLF_SMALL),
                                      ("classifier_full", classifier_oracle.DIFF_FOLEY_CLASSIFIER)])
def test_classifier_oracle_matches_reference(name, cfg):
    g = load(name)
    sd = classifier_oracle.seeded_state_dict(cfg, int(g["seed"]))
    x, t, f = (torch.from_numpy(g[k]) for k in ("x", "t", "feats"))
    with torch.no_grad():
        prob = classifier_oracle.classifier_forward(sd, cfg, x, t, f)
    assert rel_l2(prob, g["prob"]) < 1e-6
    grad = classifier_oracle.loglikelihood_grad(sd, cfg, x, t, f, 50.0)
    assert rel_l2(grad, g["grad"]) < 1e-4


def test_classifier_module_matches_oracle_and_keys():
    """The product-side classifier mirror (torch autograd, see diff_foley_b200/classifier.py) has the
    reference's parameter names and computes the same function as the oracle."""
    from diff_foley_b200.classifier import ClassifierBackboneB200
    cfg = CLF_SMALL
    m = ClassifierBackboneB200(image_size=32, in_channels=4, out_channels=1, model_channels=cfg["model_channels"],
                               attention_resolutions=list(cfg["attention_resolutions"]), num_res_blocks=1,
                               channel_mult=list(cfg["channel_mult"]), num_heads=cfg["num_heads"],
                               use_spatial_transformer=True, transformer_depth=1, context_dim=cfg["context_dim"],
                               use_checkpoint=True, legacy=False)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == dict(classifier_oracle.classifier_param_shapes(cfg))
    g = load("classifier_small")
    m.load_state_dict(classifier_oracle.seeded_state_dict(cfg, int(g["seed"])))
    x, t, f = (torch.from_numpy(g[k]) for k in ("x", "t", "feats"))
    with torch.no_grad():
        assert rel_l2(m(x, timesteps=t, context=f), g["prob"]) < 1e-6


def test_vae_decode_oracle_matches_reference():
    """oracle/vae_oracle.py vs the reference's AutoencoderKL.decode on the reference-sampled latent."""
    from oracle import vae_oracle
    g = load("vae_decode")
    sd = vae_oracle.seeded_state_dict(int(g["seed"]))
    img = vae_oracle.decode_first_stage(sd, torch.from_numpy(g["z"]))
    assert img.shape == (1, 3, 128, 512)
    assert rel_l2(img[:, 0], g["mel"]) < 1e-5


# ------------------------------------------------------------------------- CAVP encoders (a17, a18)
def cavp_inputs(g):
    B, T, HW, spec_T = (int(v) for v in g["shape"])
    gen = torch.Generator().manual_seed(int(g["seed"]) + 77)
    video = torch.rand(B, T, 3, HW, HW, generator=gen)
    spec = torch.randn(B, 128, spec_T, generator=gen)
    return video, spec


def test_cavp_oracle_matches_reference():
    from oracle import cavp_oracle
    g = load("cavp_small")
    sd = cavp_oracle.seeded_state_dict(int(g["seed"]))
    video, spec = cavp_inputs(g)
    assert rel_l2(cavp_oracle.encode_video(sd, video, normalize=False, pool=False), g["video_raw"]) < 1e-5
    assert rel_l2(cavp_oracle.encode_video(sd, video, normalize=True, pool=False), g["video_feat"]) < 1e-5
    assert rel_l2(cavp_oracle.encode_spec(sd, spec, normalize=False, pool=False), g["spec_raw"]) < 1e-5
    assert rel_l2(cavp_oracle.encode_spec(sd, spec, normalize=True, pool=False), g["spec_feat"]) < 1e-5


def test_cavp_module_keys_match_reference():
    from diff_foley_b200.cavp import CAVPInferenceB200
    from oracle import cavp_oracle
    with torch.device("meta"):
        m = CAVPInferenceB200()
    got = {k: tuple(v.shape) for k, v in m.state_dict().items() if "num_batches" not in k and k != "logit_scale"}
    assert got == dict(cavp_oracle.cavp_param_shapes())


# --------------------------------------------------------- round 2: a16, N3 (DPM-Solver / PLMS), schedule length
def test_cond_stage_oracle_matches_reference():
    """Video_Feat_Encoder_Posembed (row a16): oracle vs the reference module's output (cond_embed.npz)."""
    g = load("cond_embed")
    sd = ddim_oracle.cond_stage_seeded_state(int(g["seed"]))
    out = ddim_oracle.cond_stage(sd, torch.from_numpy(g["feats"]))
    assert out.shape == (3, 32, 768)
    assert rel_l2(out, g["out"]) < 2e-6


def test_ddim_oracle_runs_every_schedule_entry_when_S_does_not_divide_1000():
    """S = 30: c = 1000 // 30 = 33, range(0, 1000, 33) has 31 entries and the reference runs them all
    (ddim.py:197-199) -- the round-1 fused sampler stopped after S."""
    g = load("ddim_small_s30")
    c = ddim_oracle.ddim_coefficients(int(g["steps"]))
    assert len(c["timesteps"]) == len(range(0, 1000, 1000 // 30)) == 31
    sd = unet_oracle.seeded_state_dict(SMALL, int(g["seed"]))
    cond = torch.from_numpy(g["cond"])
    fn = lambda x, t, cc: unet_oracle.unet_forward(sd, SMALL, x, t, cc)
    x, _ = ddim_oracle.ddim_sample(fn, torch.from_numpy(g["x_T"]), cond, torch.zeros_like(cond), float(g["scale"]),
                                   int(g["steps"]))
    assert rel_l2(x, g["samples"]) < 2e-5


@pytest.mark.parametrize("name", ["dpm_small", "dpm_small_s10"])
def test_dpm_solver_oracle_matches_reference_sampler(name):
    from oracle import dpm_oracle
    g = load(name)
    sd = unet_oracle.seeded_state_dict(SMALL, int(g["seed"]))
    cond = torch.from_numpy(g["cond"])
    fn = lambda x, t, cc: unet_oracle.unet_forward(sd, SMALL, x, t, cc)
    x = dpm_oracle.dpm_solver_sample(fn, torch.from_numpy(g["x_T"]), cond, torch.zeros_like(cond), float(g["scale"]),
                                     int(g["steps"]))
    assert rel_l2(x, g["samples"]) < 5e-5


def test_plms_oracle_matches_reference_sampler():
    from oracle import dpm_oracle
    g = load("plms_small")
    sd = unet_oracle.seeded_state_dict(SMALL, int(g["seed"]))
    cond = torch.from_numpy(g["cond"])
    fn = lambda x, t, cc: unet_oracle.unet_forward(sd, SMALL, x, t, cc)
    x, _ = dpm_oracle.plms_sample(fn, torch.from_numpy(g["x_T"]), cond, torch.zeros_like(cond), float(g["scale"]),
                                  int(g["steps"]))
    assert rel_l2(x, g["samples"]) < 5e-5


@pytest.mark.parametrize("steps", [10, 25, 50])
def test_dpm_host_schedule_equals_oracle_schedule(steps):
    """diff_foley_b200/dpm_solver.py precomputes the per-step scalars the fused sampler consumes; they must be
    the oracle's NoiseScheduleDiscrete evaluated at the same times (host logic, no GPU)."""
    from diff_foley_b200.dpm_solver import dpm_solver_pp_2m_schedule
    from oracle import dpm_oracle
    ac = ddim_oracle.alphas_cumprod()
    s = dpm_solver_pp_2m_schedule(ac, steps)
    ns = dpm_oracle.NoiseScheduleDiscrete(ac)
    t = torch.linspace(1.0, 1e-3, steps + 1)
    assert np.allclose(s["t_cont"], t.numpy())
    assert np.allclose(s["sigma"], ns.marginal_std(t[:-1]).numpy(), rtol=1e-6)
    assert np.allclose(s["alpha"], ns.marginal_alpha(t[:-1]).numpy(), rtol=1e-6)
    assert np.allclose(s["t_input"], ns.model_input_time(t[:-1]).numpy(), rtol=1e-6)
    lam = ns.marginal_lambda(t)
    h = lam[1:] - lam[:-1]
    assert np.allclose(s["cx"], (ns.marginal_std(t[1:]) / ns.marginal_std(t[:-1])).numpy(), rtol=1e-6)
    assert np.allclose(s["a_coef"], (ns.marginal_alpha(t[1:]) * torch.expm1(-h)).numpy(), rtol=2e-5)
    assert s["order"][0] == 1 and (s["order"][1:-1] == 2).all() and s["order"][-1] == (1 if steps < 15 else 2)
    assert np.allclose(s["inv_r0"][1:], (h[1:] / h[:-1]).numpy(), rtol=1e-5) and s["inv_r0"][0] == 0


def test_dpm_host_loop_matches_reference_golden_with_oracle_unet():
    """DPMSolverSamplerB200's host loop (the path classifier guidance takes) driven by the oracle UNet on CPU:
    the sampler logic itself, independent of the kernels, against the reference sampler's latent."""
    from diff_foley_b200.dpm_solver import DPMSolverSamplerB200
    g = load("dpm_small_s10")
    sd = unet_oracle.seeded_state_dict(SMALL, int(g["seed"]))

    class Model:
        alphas_cumprod = ddim_oracle.alphas_cumprod()
        betas = torch.zeros(1)

        @staticmethod
        def apply_model(x, t, c):
            return unet_oracle.unet_forward(sd, SMALL, x, t, c)

    cond = torch.from_numpy(g["cond"])
    x, _ = DPMSolverSamplerB200(Model()).sample(int(g["steps"]), 1, (4, SMALL["latent_h"], SMALL["latent_w"]), cond,
                                                x_T=torch.from_numpy(g["x_T"]), unconditional_guidance_scale=float(g["scale"]),
                                                unconditional_conditioning=torch.zeros_like(cond))
    assert rel_l2(x, g["samples"]) < 5e-5


# -------------------------------------------------------------------------- N4: CAVP frame ingest (Pillow resample)
@pytest.mark.parametrize("H,W", [(360, 640), (100, 150), (224, 224), (480, 227), (720, 1280), (37, 53)])
def test_frames_oracle_is_bit_identical_to_pillow(H, W):
    """The frame preprocessing of Extract_CAVP_Features (demo_util.py:147-150) is Pillow arithmetic: the numpy
    restatement (oracle/frames_oracle.py) must reproduce PIL Resize + ToTensor bit for bit, up- and down-scaling."""
    from PIL import Image
    import torchvision.transforms as T
    from oracle import frames_oracle
    rng = np.random.default_rng(H * 1000 + W)
    bgr = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    ref = T.Compose([T.Resize((224, 224)), T.ToTensor()])(Image.fromarray(np.ascontiguousarray(bgr[:, :, ::-1]))).numpy()
    got = frames_oracle.preprocess_frame(bgr)
    assert got.dtype == np.float32 and got.shape == (3, 224, 224)
    assert np.array_equal(ref, got)


@pytest.mark.parametrize("n_in", [640, 360, 224, 150, 100, 53, 1920])
def test_frames_host_tables_equal_oracle_tables(n_in):
    """The product builds Pillow's fixed-point coefficient tables vectorised (diff_foley_b200/frames.py); they must
    equal the oracle's loop restatement of precompute_coeffs + normalize_coeffs_8bpc, and each row sums to ~2^22."""
    from diff_foley_b200.frames import _bilinear_tables
    from oracle import frames_oracle
    k1, b1, s1 = frames_oracle.precompute_coeffs(n_in, 224)
    k2, b2, s2 = _bilinear_tables(n_in, 224)
    assert s1 == s2 and np.array_equal(b1, b2) and np.array_equal(k1, k2)
    assert np.all(np.abs(k2.astype(np.int64).sum(1) - (1 << 22)) <= k2.shape[1])
    assert np.all(b2[:, 0] >= 0) and np.all(b2[:, 0] + b2[:, 1] <= n_in) and np.all(b2[:, 1] >= 1)
