#!/bin/bash
# One-GPU round-end evidence (run on the GPU box): the full GPU test suite, the driver's bench command, the one-GPU records of
# the other BASELINE configs, per-layer / role / timeline logs, then tools/make_profiles.sh (ncu + sanitizers).
# usage: tools/round_end.sh [round tag, default r2]        -> gpurun_out/<tag>/
set -u
tag=${1:-r2}
out=gpurun_out/$tag
mkdir -p $out/multi
timeout 420 python -m pytest tests -m gpu -q -s > $out/gpu_tests_full.log 2>&1
tail -2 $out/gpu_tests_full.log
timeout 300 python bench.py --steps 5 --warmup 3 > $out/bench_n1.json 2> $out/bench_n1.err
head -c 400 $out/bench_n1.json; echo
timeout 200 python bench.py --config 2 --steps 3 --warmup 3 --no-cpu-baseline > $out/multi/bench_n1_cfg2.json 2> $out/multi/bench_n1_cfg2.err
timeout 200 python bench.py --config 3 --steps 3 --warmup 3 --no-cpu-baseline > $out/multi/bench_n1_cfg3.json 2> $out/multi/bench_n1_cfg3.err
timeout 300 python bench.py --sweep > $out/multi/sweep_n1.json 2> $out/multi/sweep_n1.err
timeout 100 python tools/perf_layers.py 16 256 fp16 > $out/per_layer_times.log 2>&1
timeout 60 python tools/perf_layers.py 1 256 fp16 > $out/perf_b1.log 2>&1
timeout 60 python tools/perf_layers.py 4 256 fp16 > $out/perf_b4.log 2>&1
export FDSR_LIB=fastdiffsr_b200/libfdsr_prof.so
timeout 100 python tools/role_profile.py 16 256 > $out/role_cycles.log 2>&1
timeout 60 python tools/timeline.py 16 256 > $out/timeline_b16.log 2>&1
timeout 60 python tools/timeline.py 1 256 > $out/timeline_b1.log 2>&1
unset FDSR_LIB
tail -2 $out/timeline_b16.log
bash tools/make_profiles.sh $tag
