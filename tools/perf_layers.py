"""Per-op timing of the UNet at a given shape + whole-sampler throughput (GPU box tool).
python tools/perf_layers.py [B] [H] [dtype]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import fdsr_oracle as O  # noqa: E402
from fastdiffsr_b200 import Engine  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
H = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dtype = sys.argv[3] if len(sys.argv) > 3 else "fp16"
cfg = dict(O.DEFAULT_UNET)
sd = O.make_state_dict(cfg, seed=0)
tab = O.schedule_tables(O.make_beta_schedule(**O.DEFAULT_SCHEDULE))
eng = Engine(cfg, "cuda:0", dtype)
eng.load_state_dict(sd)
eng.set_schedule(tab["betas"])
cond = (torch.rand(B, 3, H, H, device="cuda") * 2 - 1)
x = torch.randn(B, 3, H, H, device="cuda")
eng.unet_forward(cond, x, 5)
torch.cuda.synchronize()
prof = eng.profile_unet(5, reps=int(os.environ.get("FDSR_PROF_REPS", "3")))
tot_ms = sum(p[1] for p in prof)
tot_fl = sum(p[2] for p in prof)
print(f"B={B} {H}x{H} {dtype}: workspace {eng.workspace_bytes() / 2**30:.2f} GiB, unet flops {eng.unet_flops() / 1e9:.1f} G")
for name, ms, fl in prof:
    print(f"{name:18s} {ms * 1000:9.1f} us  {fl / 1e9:9.2f} GF  {fl / ms / 1e9 if ms > 0 else 0:8.1f} TF/s")
print(f"sum of ops {tot_ms:.3f} ms -> {tot_fl / tot_ms / 1e9:.1f} TF/s over the UNet")
for graph in (False, True):
    eng.set_use_graph(graph)
    eng.sample(cond, seed=1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 3
    e0.record()
    for i in range(n):
        eng.sample(cond, seed=1)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"sample T=20 graph={graph}: {ms:.2f} ms / batch -> {B / ms * 1000:.1f} img/s, {ms / 20:.3f} ms per UNet step, "
          f"{eng.unet_flops() * 20 / ms / 1e9:.1f} TF/s")
