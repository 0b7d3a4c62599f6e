"""Small target for ncu: two UNet evaluations at the benchmark shape (first = warm-up)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fastdiffsr_b200 as F  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
H = int(sys.argv[2]) if len(sys.argv) > 2 else 256
n = int(sys.argv[3]) if len(sys.argv) > 3 else 2
model = sys.argv[4] if len(sys.argv) > 4 else "fastdiffsr"   # "sr3": the SR3 baseline (SelfAttention cores)
opt = F.config.default_config("sr_ddpm_test_64_256" if model == "sr3" else "sr_fastdiffsr_test_64_256")
if model == "sr3":
    opt["model"]["beta_schedule"]["val"]["n_timestep"] = 20
torch.manual_seed(0)
netG = F.define_G(opt).to("cuda")
netG.set_new_noise_schedule(opt["model"]["beta_schedule"]["val"], "cuda")
eng = netG.engine()
eng.set_use_graph(False)
cond = torch.rand(B, 3, H, H, device="cuda") * 2 - 1
x = torch.randn(B, 3, H, H, device="cuda")
for i in range(n):
    eng.unet_forward(cond, x, 10)
torch.cuda.synchronize()
print("done", eng.launch_count())
