#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 120 python -m pytest tests/test_gpu_ops.py -x -q -k "resstack" 2>&1 | tail -25 > gpurun_out/c3_resstack.txt
echo "rc=$?" >> gpurun_out/c3_resstack.txt
cat gpurun_out/c3_resstack.txt
if grep -q "passed" gpurun_out/c3_resstack.txt && ! grep -q "failed" gpurun_out/c3_resstack.txt; then
  timeout 600 python -m pytest tests/test_gpu_model.py -x -q 2>&1 | tail -15 > gpurun_out/c3_model.txt; cat gpurun_out/c3_model.txt
  for cfg in "fused:--voc-fused-mask 12" "unfused:--voc-fused-mask 0"; do
    name=${cfg%%:*}; flags=${cfg#*:}
    timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $flags > gpurun_out/c3_bench_$name.json 2> gpurun_out/c3_bench_$name.err
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c3_bench_$name.json").read().strip().splitlines()[-1])
    print("$name", round(d["ms_per_step"],2), "ms/step e2e", round(d["e2e"]["ms_per_step"],2), "launches", d["gpu_launches"], "frac", round(d["roofline"]["frac"],3), d["clocks"])
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/c3_bench_$name.err").read()[-1500:])
PY
  done
fi
