"""Summarise ncu CSV exports under profiles/<round>/ into profiles/<round>/SUMMARY.md."""
import collections
import csv
import glob
import json
import os
import sys

rd = sys.argv[1] if len(sys.argv) > 1 else "profiles/r1"


def num(v):
    try:
        return float(v.replace(",", ""))
    except Exception:
        return None


lines = [f"# ncu summary ({rd})", ""]
# ---- launch list
ll = os.path.join(rd, "launches_bench_step.csv")
if os.path.exists(ll):
    rows = [r for r in csv.reader(open(ll)) if len(r) > 5]
    hdr = rows[0]
    iK, iV, iU = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        v = num(r[iV])
        if v is None:
            continue
        if r[iU] == "ns":
            v /= 1000.0
        elif r[iU] == "ms":
            v *= 1000.0
        k = r[iK].split("(")[0].replace("void ", "").replace("fdsr::", "")
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    lines += ["## Launch list of one bench step (`ncu --metrics gpu__time_duration.sum --clock-control none`, "
              "1200 launches of `python bench.py --steps 1 --warmup 3`; cold-cache, serialised: compare shares)", "",
              "| kernel | launches | total us | share |", "|---|---:|---:|---:|"]
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{k}` | {n} | {us:.0f} | {100 * us / tot:.1f}% |")
    conv = sum(us for k, (n, us) in agg.items() if "conv_gemm" in k)
    split = ""
    bj = os.path.join(rd, "bench_n1.json")
    if os.path.exists(bj):
        b = json.load(open(bj))
        cm, st = b["roofline"]["conv_ms_per_unet_step"], b["config"]["ms_per_unet_step"]
        split = (f" (bench.py's CUDA-event split, `bench_n1.json`: conv launches {cm:.2f} ms, timed with an event pair around "
                 f"each launch, against {st:.2f} ms per UNet step of the free-running sampling loop)")
    lines += ["", f"conv_gemm_kernel share of the step: **{100 * conv / tot:.1f}%** of {tot / 1000:.1f} ms" + split, ""]
    # ---- HBM-bound helper kernels: algorithmic bytes per launch (B = 16, 256^2) / mean duration in the launch list
    px = 16 * 256 * 256
    algo = {  # kernel -> (bytes per launch, what)
        "pack_input_kernel<__half>": (px * (6 * 4 + 16 * 2), "reads cond + x_t fp32 NCHW, writes 16-channel NHWC fp16"),
        "posterior_kernel": (px * 3 * 4 * 3, "reads x_t, eps, writes x_{t-1} (fp32; z from Philox in-kernel)"),
        "res2img_kernel": (px * 3 * 4 * 3, "reads x_0, cond, writes SR (fp32)"),
        "noise_fill_kernel": (px * 3 * 4, "writes x_T (Philox normals)"),
        "slam_apply_kernel<__half>": (16 * 32 * 32 * 256 * 2 * 2, "reads + writes the (B,32,32,256) mid tensor"),
        "bicubic_v_kernel": (16 * (64 * 256 * 3 + 256 * 256 * 3 * 4), "reads the horizontally resampled u8 rows, writes cond fp32"),
    }
    peaks = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    hbm = None
    if os.path.exists(peaks):
        pk = json.load(open(peaks))
        hbm = pk.get("hbm_gbs")
    lines += ["## Elementwise / norm helper kernels: achieved HBM bandwidth (algorithmic bytes / mean launch duration from the list above; "
              "12-60 MB per launch, so launch latency is a visible part of each figure)", "",
              "| kernel | MB per launch | mean us | GB/s | traffic |", "|---|---:|---:|---:|---|"]
    for k, (nbytes, what) in algo.items():
        if k in agg and agg[k][0] > 0:
            us = agg[k][1] / agg[k][0]
            lines.append(f"| `{k}` | {nbytes / 1e6:.1f} | {us:.1f} | {nbytes / us / 1e3:.0f} | {what} |")
    if hbm:
        lines += ["", f"(measured HBM copy bandwidth of this pool, MEASURED_PEAKS.json: {hbm:.0f} GB/s; these kernels are "
                  "0.6 % of the step)", ""]
    else:
        lines.append("")
# ---- full-set captures of the elementwise kernels (tools/ncu_target_sample.py)
small = sorted(glob.glob(os.path.join(rd, "small", "*_details.csv")))
if small:
    lines += ["## Elementwise kernels under `ncu --set full` (one launch each at B = 16, 256^2)", "",
              "| kernel | duration | DRAM throughput | memory throughput | L2 hit rate | registers | achieved occupancy |",
              "|---|---:|---:|---:|---:|---:|---:|"]
    for f in small:
        rows = list(csv.reader(open(f)))
        hdr = rows[0]
        iM, iU, iV = hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
        d = {}
        for r in rows[1:]:
            key = r[iM] + ("_bw" if r[iU] in ("Gbyte/s", "Tbyte/s") else "")
            d.setdefault(key, r[iV] + " " + r[iU])
        lines.append(f"| `{os.path.basename(f).replace('_details.csv', '')}` | {d.get('Duration')} | {d.get('DRAM Throughput')} | "
                     f"{d.get('Memory Throughput_bw')} | {d.get('L2 Hit Rate')} | {d.get('Registers Per Thread')} | "
                     f"{d.get('Achieved Occupancy')} |")
    lines += ["", "These launches move 38-59 MB in 12-13 us: too short to fill the HBM pipeline (DRAM throughput 23-24 % of peak); "
              "they are 0.6 % of a sampling step, so they were left as plain vectorised float4 kernels.", ""]
# ---- per-kernel details
want = ["Duration", "SM Frequency", "Compute (SM) Throughput", "Memory Throughput", "DRAM Throughput", "L2 Hit Rate",
        "Registers Per Thread", "Dynamic Shared Memory Per Block", "Issued Warp Per Scheduler", "Executed Instructions",
        "Warp Cycles Per Issued Instruction", "Achieved Active Warps Per SM"]
names = {"1": "downs.1.block1 (64->64 3x3, N=64, 256^2)", "41": "ups.9.block2 (N=128, 128^2, GN + 1x1 residual chunks)",
         "37": "ups.7 (up2x conv 256->256 as four 2x2 phase convs, N=256, 128^2)", "47": "ups.13.block1 (128->64, N=64, 256^2)",
         "16": "downs.10.block1 (256->256 3x3, N=256, 32^2)"}
for f in sorted(glob.glob(os.path.join(rd, "conv*_details.csv"))):
    idx = os.path.basename(f).split("_")[0].replace("conv", "")
    rows = list(csv.reader(open(f)))
    hdr = rows[0]
    iS, iM, iU, iV = hdr.index("Section Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
    lines += [f"## conv launch {idx}: {names.get(idx, '')}", "", "| metric | value |", "|---|---:|"]
    for r in rows[1:]:
        if r[iM] in want:
            lines.append(f"| {r[iM]} | {r[iV]} {r[iU]} |")
    raw = f.replace("_details", "_raw")
    src = raw if os.path.exists(raw) else os.path.join("gpurun_out/profiles", os.path.basename(raw))
    if os.path.exists(src):
        rr = list(csv.reader(open(src)))
        d = dict(zip(rr[0], rr[2] if len(rr) > 2 else rr[1]))
        for k in ["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
                  "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
                  "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
                  "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"]:
            if k in d:
                unit = rr[1][rr[0].index(k)] if len(rr) > 2 else ""
                lines.append(f"| {k} | {d[k]} {unit} |")
        st = {k: num(v) for k, v in d.items() if "pcsamp_warps_issue_stalled" in k and "not_issued" not in k and num(v) is not None}
        tot = sum(st.values()) or 1
        top = ", ".join(f"{k.replace('smsp__pcsamp_warps_issue_stalled_', '')} {100 * v / tot:.0f}%"
                        for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:6])
        lines.append(f"| warp stall samples | {top} |")
    lines.append("")
open(os.path.join(rd, "SUMMARY.md"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:60]))
