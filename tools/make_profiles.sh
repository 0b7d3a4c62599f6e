#!/bin/bash
# Round profile evidence (run on the GPU box): launch list of the bench command + full-set captures
# of representative conv launches, exported as small CSV summaries.
set -u
out=gpurun_out/profiles
mkdir -p $out
# 1. every launch of one timed bench step with its device time (cold-cache, serialised)
ncu --metrics gpu__time_duration.sum --clock-control none -s 4700 -c 1200 --csv \
    --log-file $out/launches_bench_step.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/bench_under_ncu.log 2>&1
# 2. full-set captures: N=64 (downs.1.block1), N=128 (ups.9.block2), N=256 (ups.7), N=64 3-chunk (ups.13.block1)
for idx in 1 41 37 47; do
  skip=$((52 + idx))
  ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s $skip -c 1 -f -o $out/conv${idx} \
      python tools/ncu_target.py 16 256 2 > $out/conv${idx}.log 2>&1
  ncu -i $out/conv${idx}.ncu-rep --page details --csv > $out/conv${idx}_details.csv 2>/dev/null
  ncu -i $out/conv${idx}.ncu-rep --page raw --csv > $out/conv${idx}_raw.csv 2>/dev/null
done
ls -la $out
