#!/usr/bin/env python
"""Entry point with the reference's name and flags (FastDiffSR/infer.py): `python infer.py -c
config/sr_fastdiffsr_infer_x4.json` super-resolves the val folder (128 -> 512 UC-Merced shape) on the B200
path and writes `{results}/{step}_{idx}_sr.png` like infer.py:101-102."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from fastdiffsr_b200.evaluate import main  # noqa: E402

if __name__ == "__main__":
    main(default_config="config/sr_fastdiffsr_infer_x4.json", prog="infer.py")
