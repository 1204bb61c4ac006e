#!/bin/bash
# One gpurun call = parity tests + both bench arms + ncu launch list + one --set full capture (dev tool).
#   gpurun --timeout 1500 -- 'bash tools/gpu_session.sh <tag>'
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $OUT/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1
timeout 600 python bench.py > $OUT/${TAG}_bench_b200.json 2> $OUT/${TAG}_bench_b200.err
timeout 600 python bench.py --impl reference > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
timeout 300 python tools/quick_ab.py cfg3_1080p --iters 10 > $OUT/${TAG}_ab_cfg3_1080p.log 2>&1
timeout 300 python tools/quick_ab.py cfg2 --iters 10 > $OUT/${TAG}_ab_cfg2.log 2>&1
timeout 300 python tools/quick_ab.py cfg3_1080p --iters 10 --depth-batch 4 > $OUT/${TAG}_depth_batch.log 2>&1
timeout 300 python tools/quick_ab.py cfg3_1080p --iters 10 --depth-only > $OUT/${TAG}_depth_only.log 2>&1
# launch list of the bench command (shares only: ncu serialises and runs cold)
IBGS_BENCH_NOCLOCK=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
  --log-file $OUT/${TAG}_launches_bench.csv python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-train-step \
  > $OUT/${TAG}_launches_bench.log 2>&1
# full capture of the two tile renderers + sort + preprocess backward, one launch each, warm
timeout 600 ncu --set full --clock-control none --import-source on \
  -k regex:'render_backward_pairs_kernel|render_forward_kernel|preprocess_backward_kernel|preprocess_kernel|emit_instances_kernel' -s 13 -c 5 \
  -f -o $OUT/${TAG}_prof python tools/profile_one.py cfg3_1080p 2 > $OUT/${TAG}_ncu_full.log 2>&1
# full capture of the section-8f kernels (fused SSIM, Adam, batched depth preprocess), second warm launch each
timeout 600 ncu --set full --clock-control none --import-source on \
  -k regex:'ssim_forward_kernel|ssim_backward_kernel|adam_step_kernel|preprocess_depth_batch_kernel' -s 1 -c 7 \
  -f -o $OUT/${TAG}_prof_next python tools/profile_next.py cfg3_1080p > $OUT/${TAG}_ncu_next.log 2>&1
tail -3 $OUT/${TAG}_pytest_gpu.log
cat $OUT/${TAG}_bench_b200.json | cut -c1-600
cat $OUT/${TAG}_ab_cfg3_1080p.log
