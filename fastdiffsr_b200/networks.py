"""define_G: the drop-in boundary (mirrors FastDiffSR/model/networks.py:82-119)."""
from __future__ import annotations

import os

from .diffusion import GaussianDiffusion
from .unet import SR3UNet, UNet


def define_G(opt):
    model_opt = opt["model"]
    which = model_opt["which_model_G"]
    if which not in ("fastdiffsr", "ddpm"):
        raise NotImplementedError(f"which_model_G={which!r}: the B200 path implements 'fastdiffsr' and the SR3 baseline "
                                  "'ddpm' (tesr / gdp are further comparison baselines of the paper)")
    if ("norm_groups" not in model_opt["unet"]) or model_opt["unet"]["norm_groups"] is None:
        model_opt["unet"]["norm_groups"] = 32
    u = model_opt["unet"]
    model = (SR3UNet if which == "ddpm" else UNet)(in_channel=u["in_channel"], out_channel=u["out_channel"], norm_groups=u["norm_groups"],
                 inner_channel=u["inner_channel"], channel_mults=u["channel_multiplier"], attn_res=u["attn_res"],
                 res_blocks=u["res_blocks"], dropout=u["dropout"], image_size=model_opt["diffusion"]["image_size"])
    dtype = model_opt.get("compute_dtype") if hasattr(model_opt, "get") else None
    # "auto" (default): fp16 — the most accurate 16-bit mode, ~2 % faster than bf16 — until an activation is reported to
    # leave the fp16 range (a trained network's un-normalised residual stream can), then bf16 storage for good
    dtype = dtype or os.environ.get("FDSR_DTYPE", "auto")
    netG = GaussianDiffusion(model, image_size=model_opt["diffusion"]["image_size"],
                             channels=model_opt["diffusion"]["channels"], loss_type="l1",
                             conditional=model_opt["diffusion"]["conditional"],
                             schedule_opt=model_opt["beta_schedule"]["train"],
                             scale=int(256 / int(opt["datasets"]["train"]["l_resolution"])), dtype=dtype)
    if opt["phase"] == "train":
        raise NotImplementedError("phase='train' is outside the B200 sampling path")
    # The reference wraps in nn.DataParallel when distributed, and then bypasses it for sampling
    # (model/model.py:62-64); multi-GPU sampling here is one process per GPU (fastdiffsr_b200.parallel).
    return netG
