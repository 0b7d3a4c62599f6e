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
