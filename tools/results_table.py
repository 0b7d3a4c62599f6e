"""Markdown table of the committed bench records: python tools/results_table.py profiles/r2 [...files]"""
import glob
import json
import os
import sys

rd = sys.argv[1] if len(sys.argv) > 1 else "profiles/r2"
files = sorted(glob.glob(os.path.join(rd, "bench_*.json")) + glob.glob(os.path.join(rd, "multi", "bench_*.json")))
rows = []
for f in files:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception:
        continue
    if "roofline" not in d:
        continue
    r = d["roofline"]
    rows.append((d["config"]["workload"], d["n_gpus"], d["config"]["global_batch"] // d["n_gpus"], d["value"], d["e2e"]["value"],
                 d["config"]["ms_per_unet_step"], r.get("frac_executed", r["frac"]), r.get("frac_algorithmic", r["frac"]),
                 d["scaling"], d.get("shard_invariant"), (d.get("clocks") or {}).get("sm_mhz"), os.path.relpath(f, rd)))
print("| workload | GPUs | batch/GPU | images/s | end-to-end images/s | ms / UNet step | conv roofline executed / algorithmic "
      "(of 1418 TF/s sustained bf16) | scaling | shard-invariant | SM MHz | record |")
print("|---|---:|---:|---:|---:|---:|---:|---|---|---:|---|")
for w, n, b, v, e, ms, fe, fa, sc, si, mhz, f in sorted(rows, key=lambda r: (r[0].split(" T=")[0], r[1])):
    w = w.split(" T=")[0].replace("FastDiffSR ", "")
    print(f"| {w} | {n} | {b} | **{v:.1f}** | {e:.1f} | {ms:.2f} | {100 * fe:.1f} % / {100 * fa:.1f} % | {sc} | {si} | {mhz} | `{f}` |")
for f in sorted(glob.glob(os.path.join(rd, "multi", "sweep_*.json")) + glob.glob(os.path.join(rd, "sweep_*.json"))):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception:
        continue
    print(f"\nBatch sweep (configs[4]) on {d['n_gpus']} GPU(s), `{os.path.relpath(f, rd)}`: " +
          ", ".join(f"B={r['global_batch']}: {r['images_per_s']:.1f}" for r in d["sweep"]) + " images/s" +
          (f"; CPU reference path B=1: {d['cpu_baseline']['value']:.3f} images/s on {d['cpu_baseline']['cores']} cores" if d.get("cpu_baseline") else ""))
