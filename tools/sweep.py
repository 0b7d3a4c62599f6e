"""Throughput sweep over images per GPU (BASELINE configs[4], one rank of it) and the other configured shapes:
python tools/sweep.py [dtype]   ->  JSON lines {"lr","hr","batch","images_per_s","ms_per_unet_step","tflops"}.
x4 64->256 at B = 1..32, x8 32->256 at B = 8 (configs[2], 8 per rank), x4 128->512 at B = 4, 8, 16 (configs[3])."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fastdiffsr_b200 as F  # noqa: E402

dtype = sys.argv[1] if len(sys.argv) > 1 else "fp16"
opt = F.config.default_config()
opt["model"]["compute_dtype"] = dtype
torch.manual_seed(0)
netG = F.define_G(opt).to("cuda")
netG.set_new_noise_schedule(opt["model"]["beta_schedule"]["val"], "cuda")
eng = netG.engine()
cases = [(64, 256, b) for b in (1, 2, 4, 8, 16, 32)] + [(32, 256, 8)] + [(128, 512, b) for b in (4, 8, 16)]
for lr, hr, B in cases:
    g = torch.Generator().manual_seed(B)
    lr_u8 = torch.randint(0, 256, (B, lr, lr, 3), generator=g, dtype=torch.uint8).cuda()
    _, cond = eng.bicubic_u8(lr_u8, hr, hr, want_u8=False)
    for _ in range(2):
        sr = eng.sample(cond, seed=1)
    torch.cuda.synchronize()
    n = 3 if B * hr * hr >= 16 * 256 * 256 else 6
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        sr = eng.sample(cond, seed=2 + i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    assert torch.isfinite(sr).all()
    print(json.dumps({"lr": lr, "hr": hr, "batch": B, "images_per_s": B / ms * 1e3, "ms_per_unet_step": ms / 20,
                      "tflops": eng.unet_flops() * 20 / ms / 1e9, "workspace_gib": eng.workspace_bytes() / 2 ** 30}),
          flush=True)
