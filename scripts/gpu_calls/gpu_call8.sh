#!/bin/bash
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_gpu_ops.py -x -q -k "resstack" 2>&1 | tail -5 > gpurun_out/c8_ops.txt
cat gpurun_out/c8_ops.txt
if grep -q "passed" gpurun_out/c8_ops.txt && ! grep -q "failed" gpurun_out/c8_ops.txt; then
  timeout 300 python scripts/resstack_bench.py 2>&1 | tail -12 | tee gpurun_out/c8_resstack_bench.txt
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:resstack -c 1 -o gpurun_out/c8_resstack python scripts/resstack_bench.py fused > gpurun_out/c8_ncu.log 2>&1
fi
