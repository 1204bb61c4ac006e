#!/bin/bash
# the two --set full captures of the rasterizer kernels on the headline scene (dev tool)
TAG=${1:-r2}
OUT=gpurun_out
timeout 800 ncu --set full --clock-control none --import-source on \
  -k regex:'render_backward_pairs|render_forward_kernel|preprocess_backward_kernel|preprocess_kernel|emit_instances_kernel' -c 26 \
  -f -o $OUT/${TAG}_prof python tools/profile_one.py cfg3_1080p 2 > $OUT/${TAG}_ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on \
  -k regex:'sort_reorder_kernel|sort_hist_kernel|scan_' -s 60 -c 10 \
  -f -o $OUT/${TAG}_prof_sort python tools/profile_one.py cfg3_1080p 2 > $OUT/${TAG}_ncu_sort.log 2>&1
tail -2 $OUT/${TAG}_ncu_full.log
