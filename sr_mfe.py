#!/usr/bin/env python
"""Entry point with the reference's name and flags (FastDiffSR/sr_mfe.py): `python sr_mfe.py -p val -c
config/sr_fastdiffsr_test_64_256.json` runs the validation block (sr_mfe.py:258-386) on the B200 path.
`-p train` raises: training is outside this path."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from fastdiffsr_b200.evaluate import main  # noqa: E402

if __name__ == "__main__":
    main(default_config="config/sr_fastdiffsr_test_64_256.json", prog="sr_mfe.py")
