#!/bin/bash
mkdir -p gpurun_out
python scripts/resstack_bench.py 2>&1 | tail -12 | tee gpurun_out/c4_resstack_bench.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:resstack -c 2 -o gpurun_out/c4_resstack python scripts/resstack_bench.py fused > gpurun_out/c4_ncu.log 2>&1
tail -3 gpurun_out/c4_ncu.log
ls -la gpurun_out/*.ncu-rep
