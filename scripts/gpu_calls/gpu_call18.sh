#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do
timeout 300 python bench.py --steps 20 --warmup 3 --headline-only --no-cpu-full-batch > gpurun_out/c18_bench_$i.json 2>/dev/null
python - gpurun_out/c18_bench_$i.json <<PY
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(round(d["ms_per_step"],2), "e2e", round(d["e2e"]["ms_per_step"],2), d["extra"]["step_ms_spread"], d["clocks"]["sm_mhz"], d["clocks"]["sm_mhz_last_samples"])
PY
done
