#!/bin/bash
mkdir -p gpurun_out
timeout 150 python bench.py --steps 20 --warmup 3 --headline-only --no-cpu-full-batch > gpurun_out/c29_bench.json 2>/dev/null
python - <<PY
import json
d=json.loads(open("gpurun_out/c29_bench.json").read().strip().splitlines()[-1])
print(round(d["ms_per_step"],2), "e2e", round(d["e2e"]["ms_per_step"],2), d["extra"]["step_ms_spread"]["device_timed"], d["clocks"]["sm_mhz"], round(d["roofline"]["frac"],3), d["roofline"]["other_classes_ms"])
PY
