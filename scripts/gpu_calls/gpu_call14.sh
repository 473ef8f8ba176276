#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 180 python -m pytest tests/test_gpu_ops.py -x -q -k "resstack" 2>&1 | tail -5 > gpurun_out/c14_ops.txt
cat gpurun_out/c14_ops.txt
if grep -q "passed" gpurun_out/c14_ops.txt && ! grep -q "failed" gpurun_out/c14_ops.txt; then
  timeout 300 python scripts/resstack_bench.py fused 2>&1 | tail -4 | tee gpurun_out/c14_resstack_bench.txt
  timeout 600 python -m pytest tests/test_gpu_model.py -q -k "hifigan or end_to_end or synthesize" 2>&1 | tail -8
  grep real_checkpoint gpurun_out/parity_report.jsonl
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --headline-only > gpurun_out/c14_bench.json 2> gpurun_out/c14_bench.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/c14_bench.json").read().strip().splitlines()[-1])
print(round(d["ms_per_step"],2), "ms/step e2e", round(d["e2e"]["ms_per_step"],2), "frac", round(d["roofline"]["frac"],3), d["clocks"]["sm_mhz"])
PY
fi
