#!/bin/bash
# Fast iteration call: parity tests + stage timings (dev tool).   gpurun --timeout 900 -- 'bash tools/gpu_quick.sh <tag>'
TAG=${1:-q}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $OUT/${TAG}_pytest_gpu.log
tail -15 $OUT/${TAG}_pytest_gpu.log
timeout 300 python tools/quick_ab.py cfg3_1080p --iters 10 --no-ref 2>&1 | tee $OUT/${TAG}_ab_cfg3_1080p.log
timeout 300 python tools/quick_ab.py cfg2 --iters 10 --no-ref 2>&1 | tee $OUT/${TAG}_ab_cfg2.log
timeout 600 python bench.py --no-cpu-baseline > $OUT/${TAG}_bench_b200.json 2> $OUT/${TAG}_bench_b200.err
python - <<PY
import json
d=json.loads(open("$OUT/${TAG}_bench_b200.json").read().strip().splitlines()[-1])
print("views/s", d["value"], "ms/view", d["ms_per_view"], "e2e", d["e2e"]["value"] if d["e2e"] else None)
for k,v in d["stages_ms_per_launch"].items(): print(f"  {k:24s} {v:.4f}" if v else f"  {k:24s} -")
PY
tail -3 $OUT/${TAG}_bench_b200.err
