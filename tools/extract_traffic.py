"""profiles/<tag>/traffic.json from the raw pages of the full-set ncu captures (tools/make_profiles.sh writes conv<idx>_raw.csv):
DRAM bytes, tensor-pipe activity and shared-memory wavefronts per captured launch — what bench.py's roofline.traffic quotes.
python tools/extract_traffic.py profiles/r2 [output.json, default profiles/r2/traffic.json]"""
import csv
import json
import os
import sys

d = sys.argv[1] if len(sys.argv) > 1 else "profiles/r2"
LAUNCHES = {
    1: "downs.1.block1 (64->64 3x3, N=64, 256^2, CTA pairs)",
    41: "ups.9.block2 (N=128, 128^2, CTA pairs)",
    37: "ups.7 (256->256 upsample phase conv, N=256, 128^2 out, CTA pairs, two epilogue teams)",
    47: "ups.13.block1 (128->64 3x3, N=64, 256^2, CTA pairs)",
    16: "downs.10.block1 (256->256 3x3, N=256, 32^2, half tiles as CTA pairs)",
}
KEYS = ["dram__bytes_read.sum", "dram__bytes_read.sum.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum.per_second",
        "dram__bytes_write.sum", "dram__bytes_write.sum.pct_of_peak_sustained_elapsed", "dram__bytes_write.sum.per_second",
        "gpu__time_duration.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared_op_utccp.sum",
        "l1tex__data_pipe_tc_wavefronts_mem_shared_op_utccp.sum.pct_of_peak_sustained_elapsed", "launch__cluster_size",
        "launch__grid_size", "launch__registers_per_thread", "launch__registers_per_thread_allocated",
        "lts__t_sectors.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max", "sm__cycles_elapsed.max.per_second",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum"]
out = {"how": "ncu --set full --clock-control none, one launch each at B=16 256^2 (tools/make_profiles.sh); raw pages in "
              "conv*_raw.csv; extracted by tools/extract_traffic.py", "launches": {}}
for idx, name in LAUNCHES.items():
    path = os.path.join(d, f"conv{idx}_raw.csv")
    if not os.path.exists(path):
        continue
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    rec = {}
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            try:
                rec[k] = {"value": float(vals[i].replace(",", "")), "unit": units[i]}
            except ValueError:
                rec[k] = {"value": vals[i], "unit": units[i]}
    out["launches"][name] = rec
json.dump(out, open(sys.argv[2] if len(sys.argv) > 2 else os.path.join(d, "traffic.json"), "w"), indent=1)
for name, rec in out["launches"].items():
    rd, wr = rec.get("dram__bytes_read.sum", {}), rec.get("dram__bytes_write.sum", {})
    print(f"{name}: {rec.get('gpu__time_duration.sum', {}).get('value')} us, DRAM read {rd.get('value')} {rd.get('unit')}, "
          f"write {wr.get('value')} {wr.get('unit')}, tensor pipe active "
          f"{rec.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', {}).get('value')} %")
