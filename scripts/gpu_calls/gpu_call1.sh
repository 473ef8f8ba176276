#!/bin/bash
# first GPU call of round 2: f16+f8 operand split -- operator tests, model parity, A/B bench
mkdir -p gpurun_out
python -m pytest tests/test_gpu_ops.py -x -q -k "tensor_core" 2>&1 | tail -15 > gpurun_out/c1_ops.txt
python -m pytest tests/test_gpu_model.py -x -q 2>&1 | tail -15 > gpurun_out/c1_model.txt
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c1_bench_default.json 2> gpurun_out/c1_bench_default.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --voc-f8-mask 0 --fs2-f8 0 > gpurun_out/c1_bench_split3.json 2> gpurun_out/c1_bench_split3.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --voc-f8-mask 31 --fs2-f8 1 > gpurun_out/c1_bench_f8all.json 2> gpurun_out/c1_bench_f8all.err
cat gpurun_out/c1_ops.txt gpurun_out/c1_model.txt
for f in default split3 f8all; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c1_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["ms_per_step"],2), "ms/step  fs2-only", round(d["extra"]["fastspeech2_only_ms_per_step"],2), "frac", round(d["roofline"]["frac"],3), d["clocks"])
except Exception as e:
    print("$f", "FAILED", e); print(open("gpurun_out/c1_bench_$f.err").read()[-1500:])
PY
done
