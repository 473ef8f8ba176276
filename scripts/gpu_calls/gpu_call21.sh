#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 240 python -m pytest tests/test_gpu_ops.py -x -q -k "resstack" 2>&1 | tail -6 > gpurun_out/c21_ops.txt
cat gpurun_out/c21_ops.txt
if grep -q "passed" gpurun_out/c21_ops.txt && ! grep -q "failed" gpurun_out/c21_ops.txt; then
  timeout 600 python -m pytest tests/test_gpu_model.py -q -k "hifigan or end_to_end" 2>&1 | tail -6
  grep "full_size_B16" gpurun_out/parity_report.jsonl
  for cfg in "pair2:" "nopair:--voc-pair-mask 0" "pair23:--voc-pair-mask 12 --voc-fused-mask 0" ; do
    name=${cfg%%:*}; flags=${cfg#*:}
    timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --headline-only $flags > gpurun_out/c21_bench_$name.json 2> gpurun_out/c21_bench_$name.err
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c21_bench_$name.json").read().strip().splitlines()[-1])
    print("$name", round(d["ms_per_step"],2), "ms/step e2e", round(d["e2e"]["ms_per_step"],2), "fs2", round(d["extra"]["fastspeech2_only_ms_per_step"],2), "launches/step", d["gpu_launches"]//10)
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/c21_bench_$name.err").read()[-800:])
PY
  done
fi
