#!/bin/bash
# One multi-GPU bench run, launched the way the driver does (dev tool).
#   gpurun --gpus N --timeout 900 -- 'bash tools/gpu_scaling.sh N <tag>'
N=${1:-2}
TAG=${2:-sc}
OUT=gpurun_out
mkdir -p $OUT
if [ "$N" = "1" ]; then
  timeout 600 python bench.py --gpus 1 --no-cpu-baseline --no-train-step > $OUT/${TAG}_bench_${N}gpu.json 2> $OUT/${TAG}_bench_${N}gpu.err
else
  timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 10 --warmup 3 > $OUT/${TAG}_bench_${N}gpu.json 2> $OUT/${TAG}_bench_${N}gpu.err
fi
tail -2 $OUT/${TAG}_bench_${N}gpu.err
tail -1 $OUT/${TAG}_bench_${N}gpu.json | cut -c1-400
