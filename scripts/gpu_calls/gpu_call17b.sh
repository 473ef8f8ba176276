#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
( time timeout 1200 python -m pytest tests -q -m gpu --durations=5 ) > gpurun_out/c17b_tests.txt 2>&1
tail -12 gpurun_out/c17b_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
show() { python - "$1" <<PY
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], round(d["ms_per_step"],2), "ms/step e2e", round(d["e2e"]["ms_per_step"],2), "warmup", d["warmup"], "frac", round(d["roofline"]["frac"],3), d["clocks"], d["gpu_launches"])
for k,v in d.get("extra",{}).get("configs",{}).items(): print(k, {a: (round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a!="workload"})
PY
}
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/c17b_bench.json 2> gpurun_out/c17b_bench.err; show gpurun_out/c17b_bench.json
timeout 300 python bench.py --steps 20 --warmup 3 --warm-seconds 4 --headline-only > gpurun_out/c17b_bench_warm4.json 2>/dev/null; show gpurun_out/c17b_bench_warm4.json
timeout 300 python bench.py --steps 20 --warmup 3 --headline-only > gpurun_out/c17b_bench_warm1.json 2>/dev/null; show gpurun_out/c17b_bench_warm1.json
