"""Informational: the reference algorithm as plain PyTorch library kernels on the same B200 — the oracle's
functional restatement of GaussianDiffusion.super_resolution run on cuda (cuDNN convolutions, eager GroupNorm / Swish /
concat / posterior), i.e. what the reference's own code would do if it were moved to this GPU unchanged.  This is the
"library kernels to beat" figure printed next to libfdsr's; it is not part of the product and not the contract benchmark.
    python tools/bench_torch_gpu.py [--batch 16] [--hr 256] [--out gpurun_out/torch_gpu_baseline.json]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import fdsr_oracle as O  # noqa: E402
from fastdiffsr_b200 import Engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--hr", type=int, default=256)
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "torch_gpu_baseline.json"))
args = ap.parse_args()
B, H = args.batch, args.hr
dev = torch.device("cuda:0")
cfg = dict(O.DEFAULT_UNET)
sd_cpu = O.make_state_dict(cfg, seed=0)
sd = {k: v.to(dev) for k, v in sd_cpu.items()}
tab = O.schedule_tables(O.make_beta_schedule(**O.DEFAULT_SCHEDULE))
g = torch.Generator().manual_seed(1)
cond = (torch.rand(B, 3, H, H, generator=g) * 2 - 1).to(dev)
noises = torch.randn(20, B, 3, H, H, generator=g).to(dev)
torch.backends.cudnn.benchmark = True
out = {"workload": f"FastDiffSR x4 {H // 4}->{H} T=20 sampling, batch {B}, one B200", "rows": []}


def timed(fn, n=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        r = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, r


ref = None
for name, tf32, autocast in (("PyTorch eager fp32 (cuDNN, TF32 off)", False, None),
                             ("PyTorch eager fp32 storage, TF32 convolutions", True, None),
                             ("PyTorch eager autocast bf16", True, torch.bfloat16),
                             ("PyTorch eager autocast fp16", True, torch.float16)):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32

    def run():
        if autocast is None:
            return O.sample_loop(sd, cfg, tab, cond, noises)
        with torch.autocast("cuda", dtype=autocast):
            return O.sample_loop(sd, cfg, tab, cond, noises)

    ms, sr = timed(run)
    if ref is None:
        ref = sr
    out["rows"].append({"impl": name, "ms_per_batch": ms, "images_per_s": B / ms * 1e3, "ms_per_unet_step": ms / 20,
                        "rel_l2_vs_fp32": ((sr.float() - ref).norm() / ref.norm()).item()})

eng = Engine(cfg, dev, "fp16")
eng.load_state_dict(sd_cpu)
eng.set_schedule(tab["betas"])
ms, sr = timed(lambda: eng.sample(cond, noise=noises), n=5)
out["rows"].append({"impl": "libfdsr fp16 (this repository, same weights / noise)", "ms_per_batch": ms,
                    "images_per_s": B / ms * 1e3, "ms_per_unet_step": ms / 20,
                    "rel_l2_vs_fp32": ((sr - ref).norm() / ref.norm()).item()})
os.makedirs(os.path.dirname(args.out), exist_ok=True)
json.dump(out, open(args.out, "w"), indent=1)
for r in out["rows"]:
    print(f"{r['impl']:55s} {r['images_per_s']:8.1f} img/s  {r['ms_per_unet_step']:7.2f} ms/UNet step  rel-L2 vs fp32 {r['rel_l2_vs_fp32']:.2e}")
