#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 300 python -m pytest tests/test_gpu_ops.py -x -q 2>&1 | tail -8 > gpurun_out/c24_ops.txt
cat gpurun_out/c24_ops.txt
if grep -q "passed" gpurun_out/c24_ops.txt && ! grep -q "failed" gpurun_out/c24_ops.txt; then
  timeout 600 python -m pytest tests/test_gpu_model.py -x -q -k "fastspeech2 and not full_size and not config4" 2>&1 | tail -5
  timeout 900 python scripts/flip_census.py 3 2>&1 | grep -v Warn | tee gpurun_out/c24_flip_census.jsonl | cut -c1-420
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --headline-only > gpurun_out/c24_bench.json 2> gpurun_out/c24_bench.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/c24_bench.json").read().strip().splitlines()[-1])
print(round(d["ms_per_step"],2), "ms/step e2e", round(d["e2e"]["ms_per_step"],2), "fs2", round(d["extra"]["fastspeech2_only_ms_per_step"],2), "launches/step", d["gpu_launches"]//20, d["roofline"]["other_classes_launches"])
PY
fi
