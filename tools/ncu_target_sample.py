"""Small ncu target: one warm-up and one measured T=20 sampling run at the benchmark shape (all helper kernels launch)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fastdiffsr_b200 as F  # noqa: E402

opt = F.config.default_config()
torch.manual_seed(0)
netG = F.define_G(opt).to("cuda")
netG.set_new_noise_schedule(opt["model"]["beta_schedule"]["val"], "cuda")
eng = netG.engine()
eng.set_use_graph(False)
lr = torch.randint(0, 256, (16, 64, 64, 3), dtype=torch.uint8, device="cuda")
_, cond = eng.bicubic_u8(lr, 256, 256, want_u8=False)
for i in range(2):
    sr = eng.sample(cond, seed=i)
torch.cuda.synchronize()
print("done", eng.launch_count())
