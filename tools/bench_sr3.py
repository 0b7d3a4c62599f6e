"""SR3 baseline (which_model_G = "ddpm") on B200: ms per UNet step, per-op times of one UNet
evaluation (conv launches, SelfAttention cores), images/s of the full T = 1000 sampler, next to the CPU oracle
port timed on the host cores.  Not the contract benchmark (bench.py measures the FastDiffSR headline).
    python tools/bench_sr3.py [--batch 16] [--hr 256] [--T 1000] [--out gpurun_out/sr3_bench.json]"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import fastdiffsr_b200 as F  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--hr", type=int, default=256)
ap.add_argument("--T", type=int, default=1000)
ap.add_argument("--dtype", default="fp16")
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sr3_bench.json"))
ap.add_argument("--no-cpu", action="store_true")
args = ap.parse_args()

dev = torch.device("cuda:0")
opt = F.config.default_config("sr_ddpm_infer_x4" if args.hr == 512 else "sr_ddpm_test_64_256")
opt["model"]["compute_dtype"] = args.dtype
sched = dict(opt["model"]["beta_schedule"]["val"])
sched["n_timestep"] = args.T
torch.manual_seed(0)
netG = F.define_G(opt).to(dev)
netG.set_new_noise_schedule(sched, dev)
eng = netG.engine()
B, H = args.batch, args.hr
cond = (torch.rand(B, 3, H, H, generator=torch.Generator().manual_seed(1)) * 2 - 1).to(dev)
x = torch.randn(B, 3, H, H, generator=torch.Generator().manual_seed(2)).to(dev)

for _ in range(3):
    eng.unet_forward(cond, x, args.T // 2)
prof = eng.profile_unet(args.T // 2, reps=24)
conv = [(n, ms, fl) for n, ms, fl in prof if fl > 0]
attn = [(n, ms) for n, ms, fl in prof if n.endswith(".attn.core")]
conv_ms, conv_fl = sum(m for _, m, _ in conv), sum(f for _, _, f in conv)
tok = {n: None for n, _ in attn}
# attention core FLOPs: 4 * HW^2 * C per image (QK^T and PV, 2*MAC)
lv = {"downs": 4, "ups": 4, "mid": 5}
attn_rows = []
for n, ms in attn:
    hw = (H >> lv[n.split(".")[0]]) ** 2
    fl = 4.0 * hw * hw * 256 * B
    attn_rows.append({"op": n, "tokens": hw, "us": ms * 1e3, "tflops": fl / (ms * 1e-3) / 1e12})

torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
netG.super_resolution(cond[:1], False, seed=1)  # warm
l0 = eng.launch_count()
e0.record()
sr = netG.super_resolution(cond, False, seed=3)
e1.record()
torch.cuda.synchronize()
ms_total = e0.elapsed_time(e1)
out = {"workload": f"SR3 baseline x4 {H // 4}->{H}, T={args.T} sampling, batch {B}, {args.dtype}, 1 B200",
       "images_per_s": B / (ms_total / 1e3), "ms_per_batch": ms_total, "ms_per_unet_step": ms_total / args.T,
       "launches": eng.launch_count() - l0, "finite": bool(torch.isfinite(sr).all()),
       "unet_gflop_per_image": eng.unet_flops() / B / 1e9,
       "conv": {"launches": len(conv), "ms_per_unet_step": conv_ms, "tflops": conv_fl / (conv_ms * 1e-3) / 1e12},
       "attention_cores": attn_rows,
       "other_ms_per_unet_step": sum(m for n, m, f in prof if f == 0 and not n.endswith(".attn.core"))}
if not args.no_cpu:
    import fdsr_oracle as O
    torch.set_num_threads(os.cpu_count())
    sd = {k: v.detach().cpu() for k, v in netG.state_dict().items() if k.startswith("denoise_fn.")}
    x6 = torch.cat([cond[:1], x[:1]], 1).cpu()
    tt = torch.full((1,), args.T // 2, dtype=torch.long)
    O.sr3_unet_forward(sd, O.SR3_UNET, x6, tt, 256)
    t0 = time.perf_counter()
    ref = O.sr3_unet_forward(sd, O.SR3_UNET, x6, tt, 256)
    dt = time.perf_counter() - t0
    got = eng.unet_forward(cond[:1].contiguous(), x[:1].contiguous(), args.T // 2).cpu()
    out["cpu_baseline"] = {"kind": "port", "cores": os.cpu_count(), "s_per_unet_step_1_image": dt,
                           "images_per_s_extrapolated": 1.0 / (dt * args.T),
                           "sample": "one UNet evaluation of one image (1/T of a sampling run)",
                           "eps_rel_l2_gpu_vs_cpu": ((got - ref).norm() / ref.norm()).item()}
os.makedirs(os.path.dirname(args.out), exist_ok=True)
json.dump(out, open(args.out, "w"), indent=1)
print(json.dumps(out))
