"""Config loading for the reference's JSON-with-`//`-comments experiment files
(mirrors core/logger.py:21-112 for the keys define_G consumes; no experiment-directory side effects)."""
from __future__ import annotations

import json
import os
from collections import OrderedDict

_ALIASES = {
    # README.md:102 names a file that does not exist; the shipped file is infer_x4
    "sr_fastdiffsr_infer_128_512.json": "sr_fastdiffsr_infer_x4.json",
}


class NoneDict(dict):
    """core/logger.py:97-100: missing keys read as None."""

    def __missing__(self, key):
        return None


def dict_to_nonedict(opt):
    if isinstance(opt, dict):
        return NoneDict(**{k: dict_to_nonedict(v) for k, v in opt.items()})
    if isinstance(opt, list):
        return [dict_to_nonedict(v) for v in opt]
    return opt


def load_config(path: str, phase: str = "val", gpu_ids=None):
    """Parse a reference config file.  `//` starts a comment anywhere on a line (logger.py:27-32)."""
    base = os.path.basename(path)
    if not os.path.exists(path) and base in _ALIASES:
        path = os.path.join(os.path.dirname(path), _ALIASES[base])
    text = ""
    with open(path, "r") as f:
        for line in f:
            text += line.split("//")[0] + "\n"
    opt = json.loads(text, object_pairs_hook=OrderedDict)
    opt["phase"] = phase
    if gpu_ids is not None:
        opt["gpu_ids"] = [int(i) for i in str(gpu_ids).split(",")]
    gl = opt.get("gpu_ids") or []
    opt["distributed"] = len(gl) > 1
    return dict_to_nonedict(opt)


_SHAPES = {
    # name: (l_resolution, r_resolution, train dataroot, val dataroot)  — the reference's shipped configs
    "sr_fastdiffsr_test_64_256": (64, 256, "dataset/Train_64_256", "dataset/Test_Toronto_64_256"),
    "sr_fastdiffsr_train_64_256": (64, 256, "dataset/Train_64_256", "dataset/Test_Potsdam_64_256"),
    "sr_fastdiffsr_test_32_256": (32, 256, "dataset/Train_32_256", "dataset/Test_Toronto_32_256"),
    "sr_fastdiffsr_train_32_256": (32, 256, "dataset/Train_32_256", "dataset/Test_Potsdam_32_256"),
    "sr_fastdiffsr_infer_x4": (128, 512, "dataset/Train_64_256", "dataset/UCM_128_512"),
    "sr_fastdiffsr_infer_128_512": (128, 512, "dataset/Train_64_256", "dataset/UCM_128_512"),
    # the SR3 comparison baseline (which_model_G = "ddpm", config/sr_ddpm_*.json)
    "sr_ddpm_test_64_256": (64, 256, "dataset/Train_64_256", "dataset/Test_Toronto_64_256"),
    "sr_ddpm_train_64_256": (64, 256, "dataset/Train_64_256", "dataset/Test_Potsdam_64_256"),
    "sr_ddpm_test_32_256": (32, 256, "dataset/Train_32_256", "dataset/Test_Toronto_32_256"),
    "sr_ddpm_train_32_256": (32, 256, "dataset/Train_32_256", "dataset/Test_Potsdam_32_256"),
    "sr_ddpm_infer_x4": (128, 512, "dataset/Train_64_256", "dataset/UCM_128_512"),
}


def default_config(name: str = "sr_fastdiffsr_test_64_256", phase: str = "val", gpu_ids=(0,)):
    """Programmatic equivalent of the reference's config/<name>.json for every key define_G and the
    sampling path consume (config/sr_fastdiffsr_test_64_256.json:52-86): same UNet widths, T=20
    linear_cosine schedule, conditional 3-channel diffusion.  Lets the package run where the
    reference tree (and its JSON files) is absent."""
    name = name[:-5] if name.endswith(".json") else name
    if name not in _SHAPES:
        raise KeyError(f"unknown config {name!r}; known: {sorted(_SHAPES)}")
    lres, rres, train_root, val_root = _SHAPES[name]
    sr3 = name.startswith("sr_ddpm")
    sched = (dict(schedule="linear", n_timestep=1000, linear_start=1e-4, linear_end=2e-2) if sr3
             else dict(schedule="linear_cosine", n_timestep=20, linear_start=1e-6, linear_end=1e-2))
    # the x4 train/infer configs condition on 64->256 training crops; inference runs fully convolutionally
    train_l = 64 if lres == 128 else lres
    opt = {
        "name": name, "phase": phase, "gpu_ids": list(gpu_ids),
        "datasets": {
            "train": {"name": "Train", "mode": "LRHR", "dataroot": train_root, "datatype": "img",
                      "l_resolution": train_l, "r_resolution": 256, "batch_size": 4},
            "val": {"name": "Test", "mode": "LRHR", "dataroot": val_root, "datatype": "img",
                    "l_resolution": lres, "r_resolution": rres},
        },
        "model": {
            "which_model_G": "ddpm" if sr3 else "fastdiffsr", "finetune_norm": False,
            "unet": {"in_channel": 6, "out_channel": 3, "inner_channel": 64,
                     "channel_multiplier": [1, 1, 2, 2, 4, 4] if sr3 else [1, 2, 4, 4],
                     "attn_res": [16], "res_blocks": 2, "dropout": 0.2},
            "beta_schedule": {"train": dict(sched), "val": dict(sched)},
            "diffusion": {"image_size": 256, "channels": 3, "conditional": True},
        },
    }
    opt["distributed"] = len(opt["gpu_ids"]) > 1
    return dict_to_nonedict(opt)
