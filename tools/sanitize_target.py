"""Small target for compute-sanitizer: one UNet evaluation and one T=20 sampling run at 64x64 (B=1 and B=2), stream
launches (no graph), so every kernel of the path runs under the tool: conv_gemm_kernel in its single-CTA, pair, split-N,
stem-TMA and fused-posterior forms, the gate kernels, bicubic, noise, res2img, metrics."""
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
quick = len(sys.argv) > 1 and sys.argv[1] == "quick"   # one UNet evaluation at B = 2 only (a racecheck pass in about a minute)
for B in ((2,) if quick else (1, 2)):
    lr = torch.randint(0, 256, (B, 16, 16, 3), dtype=torch.uint8, device="cuda")
    _, cond = eng.bicubic_u8(lr, 64, 64, want_u8=False)
    x = torch.randn(B, 3, 64, 64, device="cuda")
    eps = eng.unet_forward(cond, x, 10)
    if quick:
        torch.cuda.synchronize()
        print("B", B, "quick ok", bool(torch.isfinite(eps).all()), eng.launch_count())
        continue
    sr, tr = eng.sample(cond, seed=3, trace=True)
    m = eng.metrics_u8(sr, cond)
    torch.cuda.synchronize()
    eng.check_overflow()
    print("B", B, "ok", bool(torch.isfinite(sr).all()), eng.launch_count())
