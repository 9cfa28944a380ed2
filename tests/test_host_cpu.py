"""CPU tests of the host-side logic and of the C-ABI library surface (no compute calls)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from diff_foley_b200 import _lib as L
from diff_foley_b200.ddim import DDIMSamplerB200, make_ddim_timesteps
from diff_foley_b200.unet import UNetModelB200
from oracle import ddim_oracle, unet_oracle
from helpers import unet_kwargs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    """The shared object loads on a CPU-only host and exports exactly what include/dfb.h declares."""
    lib = L.lib()
    header = open(os.path.join(ROOT, "include", "dfb.h")).read()
    declared = set(re.findall(r"\b(dfb_[a-z0-9_]+)\s*\(", header))
    assert declared == set(L.SIGNATURES), declared ^ set(L.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.dfb_version()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    """Without a CUDA device the product path fails loudly instead of computing on the CPU."""
    lib = L.lib()
    cfg = UNetModelB200(**unet_kwargs(unet_oracle.small_unet_cfg()))._cfg()
    h = ctypes.c_void_p()
    rc = lib.dfb_unet_create(ctypes.byref(cfg), 0, ctypes.byref(h))
    assert rc != 0 and b"no CPU fallback" in lib.dfb_last_error()
    m = UNetModelB200(**unet_kwargs(unet_oracle.small_unet_cfg()))
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.zeros(1, 4, 16, 64), torch.zeros(1, dtype=torch.long), torch.zeros(1, 32, 128))


@pytest.mark.parametrize("cfg", [unet_oracle.DIFF_FOLEY_UNET, unet_oracle.small_unet_cfg()])
def test_module_state_dict_matches_reference_keys(cfg):
    """UNetModelB200 exposes the reference's parameter names/shapes (checkpoint drop-in)."""
    with torch.device("meta"):
        m = UNetModelB200(**unet_kwargs(cfg))
    got = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    want = dict(unet_oracle.unet_param_shapes(cfg))
    assert got.keys() == want.keys()
    assert got == want


def test_unsupported_config_is_rejected():
    kw = unet_kwargs(unet_oracle.small_unet_cfg())
    kw["use_scale_shift_norm"] = True
    with pytest.raises(NotImplementedError):
        UNetModelB200(**kw)


def test_sampler_schedule_equals_oracle():
    class M:  # the attributes make_schedule reads
        num_timesteps = 1000
        alphas_cumprod = ddim_oracle.alphas_cumprod()
    s = DDIMSamplerB200(M())
    s.make_schedule(25)
    c = ddim_oracle.ddim_coefficients(25)
    assert list(make_ddim_timesteps(25, 1000)[[0, 1, -1]]) == [1, 41, 961]
    for k in ("timesteps", "sqrt_one_minus_at", "sqrt_at", "sqrt_a_prev", "dir_coef"):
        assert np.array_equal(s._steps[k], c[k]), k
    with pytest.raises(NotImplementedError):
        s.make_schedule(25, ddim_eta=1.0)


def test_vae_decoder_state_dict_matches_reference_keys():
    """AutoencoderKLDecoderB200 exposes the decode half of the reference AutoencoderKL's parameters."""
    from diff_foley_b200.vae import AutoencoderKLDecoderB200
    from oracle import vae_oracle
    with torch.device("meta"):
        m = AutoencoderKLDecoderB200()
    got = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    want = dict(vae_oracle.decoder_param_shapes())
    assert got == want
    with pytest.raises(RuntimeError, match="no CPU path"):
        AutoencoderKLDecoderB200().decode(torch.zeros(1, 4, 16, 64))
