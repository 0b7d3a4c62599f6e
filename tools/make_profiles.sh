#!/bin/bash
# Round profile evidence (run on the GPU box): launch list of the bench command, full-set captures of representative
# conv launches (exported as CSV), compute-sanitizer memcheck / racecheck of a small sampling run.
# usage: tools/make_profiles.sh [round tag, default r2]
set -u
tag=${1:-r2}
out=gpurun_out/profiles_$tag
mkdir -p $out
# 1. every launch of one timed bench step with its device time (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -s 3400 -c 1200 --csv \
    --log-file $out/launches_bench_step.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/bench_under_ncu.log 2>&1
# 2. full-set captures: N=64 pair (downs.1.block1), N=128 pair (ups.9.block2), N=256 phase conv (ups.7),
#    N=64 3-chunk pair (ups.13.block1), 32^2 split-N (downs.10.block1 = conv 16)
for idx in 1 41 37 47 16; do
  skip=$((52 + idx))
  ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s $skip -c 1 -f -o $out/conv${idx} \
      python tools/ncu_target.py 16 256 2 > $out/conv${idx}.log 2>&1
  ncu -i $out/conv${idx}.ncu-rep --page details --csv > $out/conv${idx}_details.csv 2>/dev/null
  ncu -i $out/conv${idx}.ncu-rep --page raw --csv > $out/conv${idx}_raw.csv 2>/dev/null
  rm -f $out/conv${idx}.ncu-rep
done
# 3. sanitizers on one 64^2 B=1 T=20 sampling run (fused tail, pair kernels, stem TMA) + one unet_forward
compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_target.py > $out/sanitizer_memcheck.log 2>&1
compute-sanitizer --tool racecheck --racecheck-report all --print-limit 20 python tools/sanitize_target.py > $out/sanitizer_racecheck.log 2>&1
tail -4 $out/sanitizer_memcheck.log $out/sanitizer_racecheck.log
ls -la $out
