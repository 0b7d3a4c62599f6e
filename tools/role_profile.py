"""Where do the cycles of a conv launch go?  Needs the -DFDSR_PROFILE build:
  nvcc ... -DFDSR_PROFILE -o fastdiffsr_b200/libfdsr_prof.so fastdiffsr_b200/csrc/api.cu
  FDSR_LIB=fastdiffsr_b200/libfdsr_prof.so python tools/role_profile.py [B] [H] [op names...]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fastdiffsr_b200 as F  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
H = int(sys.argv[2]) if len(sys.argv) > 2 else 256
want = sys.argv[3:] or ["downs.1.block1", "downs.1.block2", "downs.5.block1", "ups.7", "ups.9.block2", "ups.13.block1",
                        "ups.13.block2", "downs.10.block1", "final_conv", "downs.0", "downs.3"]
opt = F.config.default_config()
torch.manual_seed(0)
netG = F.define_G(opt).to("cuda")
netG.set_new_noise_schedule(opt["model"]["beta_schedule"]["val"], "cuda")
eng = netG.engine()
eng.set_use_graph(False)
cond = torch.rand(B, 3, H, H, device="cuda") * 2 - 1
x = torch.randn(B, 3, H, H, device="cuda")
eng.unet_forward(cond, x, 10)
torch.cuda.synchronize()
names = [p[0] for p in eng.profile_unet(10, reps=1)]
MMA = ["wait acc_empty", "wait a_full", "wait b_full", "issue+commit"]
EPI = ["wait acc_full", "tile hand-off", "stats flush+barriers", "TMEM load", "bias/resid math", "staging wait", "pack+STS+TMA store", "stats"]
PRO = ["tile setup (+GN table)", "issue loads (+table read)", "wait a_empty (+load latency)", "transform+STS+arrive"]
for nm in want:
    op = names.index(nm)
    eng.role_cycles(op, 10)
    cyc = eng.role_cycles(op, 10).astype(np.float64)
    act = cyc[:, 0, :].sum(axis=1) > 0
    n = int(act.sum())
    c = cyc[act].mean(axis=0)
    print(f"== {nm}: {n} CTAs; mean cycles per CTA")
    tl = cyc[act][:, 3, :]
    print("   timeline (cycles since kernel entry, median over CTAs): " + ", ".join(
        f"{lab} {np.median(tl[:, i][tl[:, i] > 0]) if (tl[:, i] > 0).any() else 0:.0f}" for i, lab in enumerate(
            ["prologue", "prev launch done", "GN table", "first patch", "first MMA", "last MMA issued", "epilogue done", "exit"])))
    for role, labels in ((0, MMA), (1, EPI), (2, PRO)):
        tot = c[role].sum()
        parts = ", ".join(f"{lab} {c[role][i]:.0f} ({100 * c[role][i] / max(tot, 1):.0f}%)" for i, lab in enumerate(labels))
        print(f"   {['MMA ', 'EPI ', 'PROD'][role]} total {tot:.0f}: {parts}")
