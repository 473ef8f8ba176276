#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
( time timeout 1200 python -m pytest tests -q -m gpu --durations=5 ) > gpurun_out/c15_tests.txt 2>&1
tail -12 gpurun_out/c15_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/c15_bench.json 2> gpurun_out/c15_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/c15_bench.json").read().strip().splitlines()[-1])
print(round(d["ms_per_step"],2), "ms/step e2e", round(d["e2e"]["ms_per_step"],2), "warmup", d["warmup"], "frac", round(d["roofline"]["frac"],3), d["clocks"])
for k,v in d["extra"]["configs"].items(): print(k, {a: (round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a!="workload"})
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-400
