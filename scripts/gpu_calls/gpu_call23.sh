#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 600 python -m pytest tests/test_gpu_model.py -q -k "hifigan or end_to_end or synthesize" 2>&1 | tail -4
grep "full_size_B16" gpurun_out/parity_report.jsonl
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --headline-only > gpurun_out/c23_bench.json 2> gpurun_out/c23_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/c23_bench.json").read().strip().splitlines()[-1])
print(round(d["ms_per_step"],2), "ms/step e2e", round(d["e2e"]["ms_per_step"],2), "frac", round(d["roofline"]["frac"],3), "launches/step", d["gpu_launches"]//20, d["clocks"]["sm_mhz"], d["clocks"]["power_w_max"])
PY
