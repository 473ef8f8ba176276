#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
( time timeout 900 python -m pytest tests -x -q -m gpu --durations=6 ) > gpurun_out/c9_tests.txt 2>&1
tail -14 gpurun_out/c9_tests.txt
for cfg in "default:" "f8all:--voc-f8-mask 31" "fused23:--voc-fused-mask 12"; do
  name=${cfg%%:*}; flags=${cfg#*:}
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $flags > gpurun_out/c9_bench_$name.json 2> gpurun_out/c9_bench_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c9_bench_$name.json").read().strip().splitlines()[-1])
    print("$name", round(d["ms_per_step"],2), "ms/step e2e", round(d["e2e"]["ms_per_step"],2), "fs2", round(d["extra"]["fastspeech2_only_ms_per_step"],2), "launches", d["gpu_launches"], "frac", round(d["roofline"]["frac"],3), d["clocks"]["sm_mhz"])
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/c9_bench_$name.err").read()[-1500:])
PY
done
