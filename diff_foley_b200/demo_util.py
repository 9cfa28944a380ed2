"""The entry points of the reference's inference/demo_util.py that the notebook calls (SURVEY 3.3 / 3.4, row
N4), on the B200 path: `Extract_CAVP_Features` (:80-173) and `load_model_from_config` (:177-193), same
constructor / call signatures and return values.

Differences that are the point of this module:
  * frames are not converted one by one on the CPU (cv2.cvtColor + PIL Resize + ToTensor + one H2D copy per
    frame, :147-150): a whole window of decoded uint8 frames is copied once and `dfb_frames_resize` produces the
    bit-identical fp32 [T,3,224,224] tensor on the GPU (diff_foley_b200/frames.py);
  * full windows of `batch_size` frames are encoded several at a time (the reference runs encode_video with
    batch 1 per window, :154-159): the windows are independent, so they form the batch dimension;
  * the config's `target:` strings that name reference classes are mapped to the B200 classes, so the
    reference's YAML files are used unchanged (`yaml.safe_load`; OmegaConf objects are accepted too).
Video decoding itself (ffmpeg re-encode to `fps`, cv2.VideoCapture) is CPU media I/O outside the hot path and is
kept as in the reference.
"""
import os
import subprocess
from pathlib import Path

import numpy as np
import torch

from .frames import preprocess_frames

# reference class -> B200 class (the plugin mechanism of SURVEY 8b: instantiate_from_config resolves `target`)
TARGET_MAP = {
    "model.cavp_model.CAVP_Inference": "diff_foley_b200.cavp.CAVPInferenceB200",
    "diff_foley.modules.diffusionmodules.openai_unetmodel.UNetModel": "diff_foley_b200.unet.UNetModelB200",
    "diff_foley.models.diffusion.ddpm.LatentDiffusion": "diff_foley_b200.ldm.LatentDiffusionB200",
}


def which_ffmpeg() -> str:
    result = subprocess.run(["which", "ffmpeg"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return result.stdout.decode("utf-8").replace("\n", "")


def reencode_video_with_diff_fps(video_path, tmp_path, extraction_fps, start_second, truncate_second) -> str:
    """demo_util.py:29-56: ffmpeg re-encode at `extraction_fps` (optionally a [start, start+truncate) cut)."""
    ff = which_ffmpeg()
    if ff == "":
        raise RuntimeError("ffmpeg not found: it is needed to re-time the video (as in the reference); pass decoded "
                           "frames to Extract_CAVP_Features.forward_frames instead")
    os.makedirs(tmp_path, exist_ok=True)
    stem = Path(video_path).stem
    if truncate_second is None:
        new_path = os.path.join(tmp_path, f"{stem}_new_fps_{extraction_fps}.mp4")
        cmd = [ff, "-hide_banner", "-loglevel", "panic", "-y", "-i", video_path, "-an", "-filter:v",
               f"fps=fps={extraction_fps}", new_path]
    else:
        new_path = os.path.join(tmp_path, f"{stem}_new_fps_{extraction_fps}_truncate_{start_second}_{truncate_second}.mp4")
        cmd = [ff, "-hide_banner", "-loglevel", "panic", "-y", "-ss", str(start_second), "-t", str(truncate_second),
               "-i", video_path, "-an", "-filter:v", f"fps=fps={extraction_fps}", new_path]
    subprocess.call(cmd)
    return new_path


def _get(cfg, key, default=None):
    return cfg.get(key, default) if hasattr(cfg, "get") else getattr(cfg, key, default)


def _plain(obj):
    """OmegaConf / dict-like -> plain python containers."""
    if hasattr(obj, "items"):
        return {k: _plain(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)) or type(obj).__name__ == "ListConfig":
        return [_plain(v) for v in obj]
    return obj


def load_config(path_or_cfg):
    if isinstance(path_or_cfg, (str, os.PathLike)):
        import yaml
        with open(path_or_cfg) as f:
            return yaml.safe_load(f)
    return _plain(path_or_cfg)


def get_obj_from_str(string):
    import importlib
    module, cls = TARGET_MAP.get(string, string).rsplit(".", 1)
    return getattr(importlib.import_module(module), cls)


def instantiate_from_config(config):
    """diff_foley/util.py:176-195 with the reference targets mapped to the B200 classes."""
    config = _plain(config)
    if "target" not in config:
        raise KeyError("Expected key `target` to instantiate.")
    target, params = config["target"], dict(config.get("params", {}) or {})
    if TARGET_MAP.get(target, target) == "diff_foley_b200.ldm.LatentDiffusionB200":
        # LatentDiffusion(first_stage_config, cond_stage_config, unet_config, ...) (ddpm.py:434-470)
        params = dict(unet_params=dict(params["unet_config"].get("params", {})),
                      cond_stage_params=dict(params.get("cond_stage_config", {}).get("params", {})) or None,
                      first_stage_params=dict(params.get("first_stage_config", {}).get("params", {})) or None,
                      **{k: params[k] for k in ("linear_start", "linear_end", "timesteps", "channels", "scale_factor")
                         if k in params})
    elif TARGET_MAP.get(target, target) == "diff_foley_b200.cavp.CAVPInferenceB200":
        params = {k: v for k, v in params.items() if k in ("video_encode", "spec_encode", "embed_dim")}
    return get_obj_from_str(target)(**params)


class Extract_CAVP_Features(torch.nn.Module):
    """demo_util.py:80-173.  `forward(video_path, start_second, truncate_second, tmp_path)` ->
    (np.float32 [T, 512] CAVP features at `fps`, path of the 21.5-fps re-encode)."""

    def __init__(self, fps=4, batch_size=2, device=None, tmp_path="./", video_shape=(224, 224), config_path=None,
                 ckpt_path=None, windows_per_call=4):
        super().__init__()
        self.fps, self.batch_size, self.device, self.tmp_path = fps, batch_size, device, tmp_path
        self.video_shape = tuple(video_shape)
        self.windows_per_call = windows_per_call
        config = load_config(config_path)
        self.stage1_model = instantiate_from_config(config["model"]).to(device)
        if ckpt_path is not None:
            self.init_first_from_ckpt(ckpt_path)
        self.stage1_model.eval()

    def init_first_from_ckpt(self, path):
        model = torch.load(path, map_location="cpu")
        if "state_dict" in list(model.keys()):
            model = model["state_dict"]
        new_model = {k.replace("module.", ""): v for k, v in model.items()}      # demo_util.py:110-113
        missing, unexpected = self.stage1_model.load_state_dict(new_model, strict=False)
        print(f"Restored from {path} with {len(missing)} missing and {len(unexpected)} unexpected keys")

    @torch.no_grad()
    def forward_frames(self, frames_bgr_u8):
        """frames: uint8 [N,H,W,3] in cv2's BGR order (what cap.read() yields) -> np.float32 [N, 512].
        Frames are grouped into windows of `batch_size` exactly like the reference loop (:153-166: full
        windows, then one shorter window with the remainder); full windows are encoded together."""
        frames = np.ascontiguousarray(frames_bgr_u8)
        n, bs = frames.shape[0], self.batch_size
        n_full = n // bs
        feats = []
        for w0 in range(0, n_full, self.windows_per_call):
            w1 = min(w0 + self.windows_per_call, n_full)
            x = preprocess_frames(frames[w0 * bs:w1 * bs], self.video_shape, bgr=True, device=self.device)
            x = x.view(w1 - w0, bs, 3, *self.video_shape)
            f = self.stage1_model.encode_video(x, normalize=True, pool=False)          # [W, bs, 512]
            feats.append(f.reshape(-1, f.shape[-1]).float().cpu().numpy())
        if n % bs:
            x = preprocess_frames(frames[n_full * bs:], self.video_shape, bgr=True, device=self.device).unsqueeze(0)
            f = self.stage1_model.encode_video(x, normalize=True, pool=False)
            feats.append(f.reshape(-1, f.shape[-1]).float().cpu().numpy())
        return np.concatenate(feats) if feats else np.zeros((0, 512), np.float32)

    @torch.no_grad()
    def forward(self, video_path, start_second=None, truncate_second=None, tmp_path="./tmp_folder"):
        import cv2
        self.tmp_path = tmp_path
        low = reencode_video_with_diff_fps(video_path, self.tmp_path, self.fps, start_second, truncate_second)
        high = reencode_video_with_diff_fps(video_path, self.tmp_path, 21.5, start_second, truncate_second)
        cap = cv2.VideoCapture(low)
        frames = []
        while cap.isOpened():
            ok, bgr = cap.read()
            if not ok:
                if frames:
                    break
                if cap.get(cv2.CAP_PROP_POS_FRAMES) >= cap.get(cv2.CAP_PROP_FRAME_COUNT):
                    break
                continue        # (the reference skips leading unreadable frames, :141-144)
            frames.append(bgr)
        cap.release()
        return self.forward_frames(np.stack(frames)), high


def load_model_from_config(config, ckpt, verbose=False):
    """demo_util.py:177-193: build the model named by `config.model`, load `ckpt['state_dict']` non-strictly,
    move to the GPU, eval()."""
    print(f"Loading model from {ckpt}")
    pl_sd = torch.load(ckpt, map_location="cpu")
    if "global_step" in pl_sd:
        print(f"Global Step: {pl_sd['global_step']}")
    cfg = load_config(config)
    model = instantiate_from_config(cfg["model"])
    m, u = model.load_state_dict(pl_sd["state_dict"], strict=False)
    if verbose:
        print(f"missing {len(m)} unexpected {len(u)}")
    model.cuda()
    model.eval()
    return model
