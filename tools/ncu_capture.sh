#!/bin/bash
# Full-set ncu capture of selected conv launches of one UNet step (run on the GPU box).
# usage: tools/ncu_capture.sh <tag> <conv index> [<conv index> ...]   (index = position among the 52 conv launches)
set -u
tag=$1; shift
mkdir -p gpurun_out
for idx in "$@"; do
  skip=$((52 + idx))
  out=gpurun_out/${tag}_conv${idx}
  ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s $skip -c 1 -f -o $out \
      python tools/ncu_target.py 16 256 2 > ${out}.log 2>&1
  ncu -i ${out}.ncu-rep --page raw --csv > ${out}_raw.csv 2>/dev/null
  ncu -i ${out}.ncu-rep --page details --csv > ${out}_details.csv 2>/dev/null
  ls -la ${out}.ncu-rep
done
