"""Tensor-pipe idle time BETWEEN consecutive conv launches inside a running UNet evaluation.  Needs the -DFDSR_PROFILE build:
  nvcc ... -DFDSR_PROFILE -o fastdiffsr_b200/libfdsr_prof.so fastdiffsr_b200/csrc/api.cu
  FDSR_LIB=fastdiffsr_b200/libfdsr_prof.so python tools/timeline.py [B] [H] [raw.npy]

Every CTA stamps its SM's clock at: kernel entry, prologue done, previous launch complete (griddepcontrol.wait returned),
GroupNorm table built, first patch landed, first MMA issued, last MMA issued, last accumulator complete, epilogue done, exit.  Launches k and k+1 that
ran on the same SM share the clock, so the chain  last accumulator complete (k) -> epilogue done (k) -> exit (k) -> entry (k+1)
-> prologue done -> dependency released (k+1)
-> table (k+1) -> patch (k+1) -> first MMA (k+1)  is read per SM and reported as medians over SMs, in microseconds at the
SM clock measured from %globaltimer over the evaluation."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fastdiffsr_b200 as F  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
H = int(sys.argv[2]) if len(sys.argv) > 2 else 256
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
tl = eng.timeline(10, reps=4)  # (ops, sms, 5, 8)
nops = tl.shape[0]

# SM clock in MHz from the evaluation's first and last conv: cycles on one SM / globaltimer ns
conv_ops = [i for i in range(nops) if tl[i, :, 4, 1].any()]
first, last = conv_ops[0], conv_ops[-1]
mhz = []
for sm in range(tl.shape[1]):
    a = tl[first][tl[first, :, 4, 0] == sm]
    b = tl[last][tl[last, :, 4, 0] == sm]
    a = a[a[:, 4, 1] > 0]
    b = b[b[:, 4, 1] > 0]
    if len(a) and len(b) and b[0, 4, 2] > a[0, 4, 2]:
        mhz.append((b[0, 4, 1] - a[0, 4, 1]) / (b[0, 4, 2] - a[0, 4, 2]) * 1e3)
clk = float(np.median(mhz))
span_us = (tl[last, :, 4, 2].max() - tl[first][tl[first, :, 4, 2] > 0][:, 4, 2].min()) / 1e3
print(f"B={B} {H}x{H}: SM clock {clk:.0f} MHz over the evaluation; first conv entry -> last conv entry {span_us:.1f} us")


def by_sm(op, leaders_only):
    """{sm: (entry clock, timeline stamps as absolute clocks, drain clock)} of the CTAs of op (leaders_only: those that
    issued MMAs — the peer CTA of a pair does not)."""
    out = {}
    for cta in tl[op]:
        if cta[4, 1] == 0 or (leaders_only and cta[3, 4] == 0):
            continue
        out[int(cta[4, 0])] = (int(cta[4, 1]), cta[3].astype(np.int64) + int(cta[4, 1]), int(cta[4, 1] + cta[4, 3]))
    return out


us = lambda cyc: cyc / clk
hdr = ("op k -> op k+1", "drained->epi", "epi->exit", "exit->entry", "entry->prol", "prol->dep", "dep->table", "table->patch",
       "patch->MMA", "idle us", "busy us")
print("%-34s %12s %10s %11s %11s %9s %10s %12s %10s %8s %8s" % hdr)
if len(sys.argv) > 3:
    np.save(sys.argv[3], tl)
tot_idle = tot_busy = 0.0
rows = []
for a, b in zip(conv_ops[:-1], conv_ops[1:]):
    A, Bm = by_sm(a, False), by_sm(b, True)
    sms = sorted(set(A) & set(Bm))
    if not sms:
        continue
    seg = np.zeros((len(sms), 9))
    busy = np.zeros(len(sms))
    for j, sm in enumerate(sms):
        _, ta, a_drain = A[sm]
        eb, tb, b_drain = Bm[sm]
        last_mma, epi, ex = a_drain, ta[6], ta[7]  # (pipe idle from the moment the last accumulator is complete)
        dep = tb[1]
        # layers without GroupNorm have no table / patch stamps (zero offsets): collapse those segments
        table = tb[2] if tb[2] > eb else dep
        patch = tb[3] if tb[3] > eb else table
        mma = tb[4]
        seg[j] = [epi - last_mma, ex - epi, eb - ex, tb[0] - eb, dep - tb[0], table - dep, patch - table, mma - patch,
                  mma - last_mma]
        busy[j] = b_drain - tb[4]
    med = np.median(seg, axis=0)
    idle = us(med[8])
    gap_ops = b - a - 1  # non-conv ops in between (the CLAM / SLAM gate kernels)
    tag = f"{names[a]} -> {names[b]}" + (" (+gates)" if gap_ops else "")
    print("%-34s %12.2f %10.2f %11.2f %11.2f %9.2f %10.2f %12.2f %10.2f %8.2f %8.2f" % (
        (tag[:34],) + tuple(us(med[i]) for i in range(8)) + (idle, us(np.median(busy)))))
    tot_idle += idle
    tot_busy += us(np.median(busy))
    rows.append(med)
# inside dep -> table (producer thread 0 of every CTA): launch dependency seen -> statistics arrived -> entries written -> barrier
pd = []
for op in conv_ops:
    c = tl[op]
    ok = (c[:, 4, 4] > 0) & (c[:, 4, 5] > 0) & (c[:, 4, 6] > 0) & (c[:, 3, 2] > 0)
    if ok.any():
        pd.append(np.median(np.stack([c[ok, 4, 5] - c[ok, 4, 4], c[ok, 4, 6] - c[ok, 4, 5], c[ok, 3, 2] - c[ok, 4, 6],
                                      c[ok, 4, 4] - c[ok, 3, 1], c[ok, 4, 7] - c[ok, 4, 4]], 1), axis=0))
if pd:
    m4 = np.median(np.array(pd), axis=0)
    print(f"GroupNorm table (median over layers): dependency seen by the producers {us(m4[3]):+.2f} us after warp 0, "
          f"statistics arrived +{us(m4[0]):.2f}, entries written +{us(m4[1]):.2f}, barrier passed +{us(m4[2]):.2f} us; "
          f"a plain 16-byte load of the patch's first pixel issued at the dependency had arrived by +{us(m4[4]):.2f} us")
print(f"sum over {len(rows)} layer boundaries: tensor pipe idle {tot_idle:.1f} us, issuing {tot_busy:.1f} us "
      f"({100 * tot_idle / (tot_idle + tot_busy):.1f} % of the evaluation idle between layers)")
m = np.median(np.array(rows), axis=0)
print("median boundary: " + ", ".join(f"{lab} {us(v):.2f}" for lab, v in zip(hdr[1:10], m)) + " us")
