# usage: bash tools/run_variants.sh <scene> <variant names...>   ("default" = the in-tree library)
SCENE=$1; shift
for v in "$@"; do
  if [ "$v" = default ]; then python tools/stage_times.py $SCENE --no-ref 2>&1 | tail -1 | sed 's/preprocess=.*render_forward/render_forward/';
  else IBGS_B200_LIB=ibgs_b200/_lib/variants/$v/libibgs_b200.so python tools/stage_times.py $SCENE --no-ref 2>&1 | tail -1 | sed 's/preprocess=.*render_forward/render_forward/'; fi
done
