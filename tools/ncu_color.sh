TAG=r2_l
OUT=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on \
  -k regex:'color_features' -s 4 -c 2 \
  -f -o $OUT/${TAG}_prof_color python tools/profile_color.py > $OUT/${TAG}_ncu_color.log 2>&1
tail -3 $OUT/${TAG}_ncu_color.log
