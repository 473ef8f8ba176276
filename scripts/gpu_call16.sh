#!/bin/bash
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_ops.py -x -q -k "attention" 2>&1 | tail -15 > gpurun_out/c16_ops.txt
cat gpurun_out/c16_ops.txt
if grep -q "passed" gpurun_out/c16_ops.txt && ! grep -q "failed" gpurun_out/c16_ops.txt; then
  timeout 600 python -m pytest tests/test_gpu_model.py -x -q -k "fastspeech2 and not full_size and not config4" 2>&1 | tail -6
  for cfg in "fusedatt:" ; do
    timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --headline-only > gpurun_out/c16_bench.json 2> gpurun_out/c16_bench.err
    python - <<PY
import json
d=json.loads(open("gpurun_out/c16_bench.json").read().strip().splitlines()[-1])
print(round(d["ms_per_step"],2), "ms/step e2e", round(d["e2e"]["ms_per_step"],2), "fs2", round(d["extra"]["fastspeech2_only_ms_per_step"],2), "att ms", d["roofline"]["other_classes_ms"]["attention"], d["roofline"]["other_classes_launches"])
PY
  done
fi
