#!/bin/bash
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_gpu_ops.py -x -q -k "resstack or test_conv1d" 2>&1 | tail -5 > gpurun_out/c18_ops.txt
cat gpurun_out/c18_ops.txt
if grep -q "passed" gpurun_out/c18_ops.txt && ! grep -q "failed" gpurun_out/c18_ops.txt; then
  timeout 300 python scripts/resstack_bench.py 2>&1 | tail -10 | tee gpurun_out/c18_resstack_bench.txt
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --headline-only --voc-fused-mask 12 > gpurun_out/c18_bench_fused23.json 2> gpurun_out/c18_bench.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/c18_bench_fused23.json").read().strip().splitlines()[-1])
print("fused23", round(d["ms_per_step"],2), "ms/step e2e", round(d["e2e"]["ms_per_step"],2), "fs2", round(d["extra"]["fastspeech2_only_ms_per_step"],2), d["roofline"]["other_classes_ms"])
PY
fi
