#!/bin/bash
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_gpu_ops.py -x -q -k "resstack or f8_split" 2>&1 | tail -25 > gpurun_out/c6_ops.txt
echo "rc=$?" >> gpurun_out/c6_ops.txt
cat gpurun_out/c6_ops.txt
if grep -q "passed" gpurun_out/c6_ops.txt && ! grep -q "failed" gpurun_out/c6_ops.txt; then
  timeout 300 python scripts/resstack_bench.py 2>&1 | tail -12 | tee gpurun_out/c6_resstack_bench.txt
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fs2 -s 500 -c 260 --csv --log-file gpurun_out/c6_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/c6_b.log 2>&1
  tail -1 gpurun_out/c6_b.log | cut -c1-200
fi
