#!/bin/bash
mkdir -p gpurun_out
run() { tag=$1; shift
timeout 300 python bench.py --steps 40 --warmup 3 --headline-only --no-cpu-baseline "$@" > gpurun_out/c21b_$tag.json 2>/dev/null
python - gpurun_out/c21b_$tag.json $tag <<PY
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], round(d["ms_per_step"],2), "e2e", round(d["e2e"]["ms_per_step"],2), "warmup", d["warmup"], d["extra"]["step_ms_spread"]["device_timed"], "e2e", d["extra"]["step_ms_spread"]["e2e"])
PY
}
run prime0a --prime 0
run prime2a
run prime0b --prime 0
run prime2b
