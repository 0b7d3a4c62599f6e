"""Layer-by-layer comparison of one libfdsr evaluation of the SR3 baseline UNet (which_model_G = "ddpm")
against the CPU oracle (bring-up tool).  Run on a GPU box:
    python tools/debug_sr3.py [H] [B] [dtype] [image_size] [t]        (FDSR_ATTN_REF=1: CUDA-core attention)"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import fdsr_oracle as O  # noqa: E402
from fastdiffsr_b200 import Engine  # noqa: E402

H = int(sys.argv[1]) if len(sys.argv) > 1 else 128
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dtype = sys.argv[3] if len(sys.argv) > 3 else "fp16"
image_size = int(sys.argv[4]) if len(sys.argv) > 4 else 256
t = int(sys.argv[5]) if len(sys.argv) > 5 else 11

cfg = dict(O.SR3_UNET)
sd = O.make_state_dict(cfg, seed=3, gn_jitter=0.2, spec=O.sr3_state_dict_spec(cfg, image_size))
tab = O.schedule_tables(O.make_beta_schedule(schedule="linear", n_timestep=20, linear_start=1e-4, linear_end=0.2))
g = torch.Generator().manual_seed(3)
cond = torch.rand(B, 3, H, H, generator=g) * 2 - 1
x = torch.randn(B, 3, H, H, generator=g)
taps = {}
t0 = time.time()
eps_ref = O.sr3_unet_forward(sd, cfg, torch.cat([cond, x], 1), torch.full((B,), t, dtype=torch.long), image_size, taps=taps)
print(f"oracle forward {time.time() - t0:.2f}s")

eng = Engine(dict(cfg, model="ddpm", image_size=image_size), "cuda:0", dtype)
eng.load_state_dict(sd)
eng.set_schedule(tab["betas"])
eng.set_use_graph(False)
eps = eng.unet_forward(cond.cuda(), x.cuda(), t)
torch.cuda.synchronize()
print("launches:", eng.launch_count(), "flops/forward: %.3f G" % (eng.unet_flops() / 1e9))


def rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


worst = 0.0
for name in eng.tensor_names():
    if name not in taps:
        continue
    ref = taps[name]
    got = eng.read_tensor(name, B, ref.numel()).cpu()
    r = rel(got, ref)
    worst = max(worst, r)
    print(f"{name:16s} C={ref.shape[1]:4d} {ref.shape[2]:4d}x{ref.shape[3]:<4d} rel-L2 {r:.3e}  max|d| {(got - ref).abs().max():.3e}  |ref|max {ref.abs().max():.2f}")
r = rel(eps.cpu(), eps_ref)
print(f"eps rel-L2 {r:.3e}   (worst layer {worst:.3e})")
torch.cuda.synchronize()
t0 = time.time()
for _ in range(5):
    eng.unet_forward(cond.cuda(), x.cuda(), t)
torch.cuda.synchronize()
print(f"unet_forward wall {1000 * (time.time() - t0) / 5:.2f} ms (B={B}, {H}x{H}, stream launches)")
sys.exit(0 if r < 1e-2 else 1)
