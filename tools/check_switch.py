"""Does an environment switch change the result?  python tools/check_switch.py NAME=VALUE [B] [H]
One UNet evaluation and one T=20 sampling run with and without the switch (read when the context is created); prints
whether eps / SR are bit-identical, else their relative L2 difference."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import fdsr_oracle as O  # noqa: E402
from fastdiffsr_b200 import Engine  # noqa: E402

name, _, val = sys.argv[1].partition("=")
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
H = int(sys.argv[3]) if len(sys.argv) > 3 else 128
cfg = dict(O.DEFAULT_UNET)
sd = O.make_state_dict(cfg, seed=0, gn_jitter=0.2)
betas = O.schedule_tables(O.make_beta_schedule(**O.DEFAULT_SCHEDULE))["betas"]


def run(env):
    os.environ.update(env)
    try:
        e = Engine(cfg, "cuda:0", "fp16")
    finally:
        for k in env:
            os.environ.pop(k, None)
    e.load_state_dict(sd)
    e.set_schedule(betas)
    g = torch.Generator().manual_seed(3)
    cond = (torch.rand(B, 3, H, H, generator=g) * 2 - 1).cuda()
    x = torch.randn(B, 3, H, H, generator=g).cuda()
    eps = e.unet_forward(cond, x, 7).clone()
    sr = e.sample(cond, seed=11).clone()
    e.check_overflow()
    e.close()
    return eps, sr


a, b = run({}), run({name: val})
for lab, u, v in (("eps", a[0], b[0]), ("SR (T=20)", a[1], b[1])):
    same = torch.equal(u, v)
    rel = ((u.double() - v.double()).norm() / u.double().norm()).item()
    print(f"{name}={val} B={B} {H}x{H}: {lab} {'bit-identical' if same else f'DIFFERS, rel-L2 {rel:.3e}'}")
