#!/bin/bash
mkdir -p gpurun_out
run() { tag=$1; shift
timeout 300 python bench.py --steps 60 --warmup 3 --headline-only --no-cpu-baseline "$@" > gpurun_out/c20b_$tag.json 2>/dev/null
python - gpurun_out/c20b_$tag.json $tag <<PY
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
c=d["clocks"] or {}
print(sys.argv[2], round(d["ms_per_step"],2), "e2e", round(d["e2e"]["ms_per_step"],2), d["extra"]["step_ms_spread"]["device_timed"], "e2e max", d["extra"]["step_ms_spread"]["e2e"]["max"], c.get("sm_mhz"), c.get("query_ms_max"), c.get("samples"))
PY
}
run all100
run off --sampler-period-ms 0
run clocks100 --sampler-queries clocks
run all250 --sampler-period-ms 250
run off2 --sampler-period-ms 0
run all100b
