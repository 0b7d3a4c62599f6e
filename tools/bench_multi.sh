#!/bin/bash
# All BASELINE.json configs on one 8-GPU box (run under `gpurun --gpus 8`): configs[1] weak scaling at 8, configs[2] (x8, global
# batch 64), configs[3] (128->512, global batch 32) at 2 / 4 / 8, configs[4] batch sweep on 8.  Records: gpurun_out/r2/multi/
set -u
out=gpurun_out/r2/multi
mkdir -p $out
port=29600
run() {  # n, name, args...
  n=$1; name=$2; shift; shift
  port=$((port + 1))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $n "$@" > $out/$name.json 2> $out/$name.err
  echo "$name: $(head -c 330 $out/$name.json)"
}
run 8 bench_n8_cfg1 --steps 4 --warmup 3
run 8 bench_n8_cfg2 --config 2 --steps 4 --warmup 3
run 8 bench_n8_cfg3 --config 3 --steps 3 --warmup 3
run 4 bench_n4_cfg3 --config 3 --steps 3 --warmup 3
run 2 bench_n2_cfg3 --config 3 --steps 3 --warmup 3
run 8 sweep_n8 --sweep --no-cpu-baseline
timeout 300 python bench.py --config 2 --steps 3 --warmup 3 --no-cpu-baseline > $out/bench_n1_cfg2.json 2> $out/bench_n1_cfg2.err
timeout 300 python bench.py --config 3 --steps 3 --warmup 3 --no-cpu-baseline > $out/bench_n1_cfg3.json 2> $out/bench_n1_cfg3.err
timeout 300 python bench.py --sweep > $out/sweep_n1.json 2> $out/sweep_n1.err
echo "n1 cfg2: $(head -c 200 $out/bench_n1_cfg2.json)"; echo "n1 cfg3: $(head -c 200 $out/bench_n1_cfg3.json)"
tail -c 600 $out/sweep_n8.json
