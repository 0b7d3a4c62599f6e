import sys, os, torch
sys.path.insert(0, os.getcwd())
import fastdiffsr_b200 as F
opt = F.config.default_config("sr_ddpm_test_64_256")
opt["model"]["beta_schedule"]["val"]["n_timestep"] = 20
torch.manual_seed(0)
netG = F.define_G(opt).to("cuda"); netG.set_new_noise_schedule(opt["model"]["beta_schedule"]["val"], "cuda")
eng = netG.engine()
B, H = 16, 256
cond = torch.rand(B, 3, H, H, device="cuda") * 2 - 1; x = torch.randn(B, 3, H, H, device="cuda")
for _ in range(3): eng.unet_forward(cond, x, 10)
prof = eng.profile_unet(10, reps=12)
tot = sum(p[1] for p in prof)
for n, ms, fl in prof:
    print(f"{n:22s} {ms*1000:8.1f} us {fl/1e9:8.2f} GF {fl/ms/1e9 if ms>0 else 0:8.1f} TF/s")
print(f"sum {tot:.3f} ms")
