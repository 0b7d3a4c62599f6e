"""A/B timing of library builds on one box in one session (GPU clocks drift by a few % between boxes / runs, so
small changes can only be judged back to back): python tools/ab_perf.py libA.so libB.so[@ENV=VAL,...] [...] [--reps 3] [--B 16] [--H 256]
Prints per build the mean over rounds of (sum of per-op times of one UNet evaluation, T=20 sampling time)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
args = sys.argv[1:]
reps, B, H, dtype = 3, "16", "256", "fp16"
libs = []
i = 0
while i < len(args):
    if args[i] == "--reps":
        reps = int(args[i + 1]); i += 2
    elif args[i] == "--B":
        B = args[i + 1]; i += 2
    elif args[i] == "--H":
        H = args[i + 1]; i += 2
    elif args[i] == "--dtype":
        dtype = args[i + 1]; i += 2
    else:
        libs.append(args[i]); i += 1
res = {l: [] for l in libs}
for r in range(reps):
    for l in libs:
        path, _, envs = l.partition("@")
        env = dict(os.environ, FDSR_LIB=os.path.abspath(path))
        for kv in filter(None, envs.split(",")):
            k, _, v = kv.partition("=")
            env[k] = v
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "perf_layers.py"), B, H, dtype], env=env,
                             capture_output=True, text=True).stdout
        ops = [x for x in out.splitlines() if x.startswith("sum of ops")]
        smp = [x for x in out.splitlines() if x.startswith("sample T=20 graph=True")]
        if ops and smp:
            res[l].append((float(ops[0].split()[3]), float(smp[0].split()[3])))
        if r == reps - 1:
            open(os.path.join(ROOT, "gpurun_out", "ab_" + os.path.basename(l).replace("=", "") + ".log"), "w").write(out)
for l in libs:
    v = res[l]
    if v:
        print(f"{os.path.basename(l):28s} sum of ops {sum(a for a, _ in v) / len(v):.3f} ms   sampler {sum(b for _, b in v) / len(v):.2f} ms/batch   "
              f"rounds: {' '.join(f'{a:.3f}/{b:.2f}' for a, b in v)}")
