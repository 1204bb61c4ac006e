#!/bin/bash
# One gpurun call = parity tests + smoke + both bench arms + A/B logs + ncu launch list + --set full captures (dev tool).
#   gpurun --timeout 1700 -- 'bash tools/gpu_session_r2.sh <tag>'      then   python tools/make_profiles.py <tag> r2 <scaling tag>
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
timeout 700 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $OUT/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1
( time timeout 700 python bench.py > $OUT/${TAG}_bench_b200.json 2> $OUT/${TAG}_bench_b200.err ) 2> $OUT/${TAG}_bench_b200.time
cp $OUT/${TAG}_bench_b200.json $OUT/${TAG}_bench_1gpu.json
( time timeout 700 python bench.py --impl reference > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err ) 2> $OUT/${TAG}_bench_ref.time
timeout 300 python tools/quick_ab.py cfg3_1080p --iters 10 > $OUT/${TAG}_ab_cfg3_1080p.log 2>&1
timeout 300 python tools/quick_ab.py cfg2 --iters 10 > $OUT/${TAG}_ab_cfg2.log 2>&1
timeout 300 python tools/quick_ab.py cfg3_1080p --iters 10 --depth-batch 4 > $OUT/${TAG}_depth_batch.log 2>&1
timeout 300 python tools/quick_ab.py cfg3_1080p --iters 10 --depth-only > $OUT/${TAG}_depth_only.log 2>&1
timeout 250 python tools/colorfeat_bench.py 2>&1 | grep "fwd+bwd\|color_features" > $OUT/${TAG}_colorfeat_bench.log
if [ "$2" != "noncu" ]; then
IBGS_BENCH_NOCLOCK=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
  --log-file $OUT/${TAG}_launches_bench.csv python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-train-step --no-extras \
  > $OUT/${TAG}_launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on \
  -k regex:'render_backward_pairs|render_forward_kernel|preprocess_backward_kernel|preprocess_kernel|emit_instances_kernel|sort_reorder_kernel' -s 20 -c 14 \
  -f -o $OUT/${TAG}_prof python tools/profile_one.py cfg3_1080p 2 > $OUT/${TAG}_ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on \
  -k regex:'ssim_forward_kernel|ssim_backward_kernel|adam_step_kernel|preprocess_depth_batch_kernel' -s 1 -c 7 \
  -f -o $OUT/${TAG}_prof_next python tools/profile_next.py cfg3_1080p > $OUT/${TAG}_ncu_next.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on \
  -k regex:'color_features|relu_bias_backward|maxpool2|upsample' -s 12 -c 12 \
  -f -o $OUT/${TAG}_prof_color python tools/profile_color.py > $OUT/${TAG}_ncu_color.log 2>&1
fi
tail -3 $OUT/${TAG}_pytest_gpu.log
tail -2 $OUT/${TAG}_smoke.log
cut -c1-300 $OUT/${TAG}_bench_b200.json
