#!/bin/bash
# usage: bash scripts/multi_gpu_bench.sh N [extra]   (on a box with N GPUs; results under gpurun_out/)
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR scripts/multi_gpu_render_check.py > gpurun_out/r2_mg${N}_check.log 2>&1; tail -1 gpurun_out/r2_mg${N}_check.log
$TR bench.py --gpus $N --steps 200 --warmup 20 > gpurun_out/r2_bench_weak_${N}gpu.json 2> gpurun_out/r2_bench_weak_${N}gpu.err
$TR bench.py --gpus $N --steps 200 --warmup 20 --scaling strong > gpurun_out/r2_bench_strong_${N}gpu.json 2> gpurun_out/r2_bench_strong_${N}gpu.err
if [ "$2" = extra ]; then
  $TR scripts/extra_configs.py hugs render --steps 100 > gpurun_out/r2_extra_${N}gpu.jsonl 2> gpurun_out/r2_extra_${N}gpu.err
  $TR bench.py --gpus $N --config Aprime --steps 30 --warmup 5 > gpurun_out/r2_bench_Aprime_${N}gpu.json 2> gpurun_out/r2_bench_Aprime_${N}gpu.err
fi
