#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:resstack -c 1 -o gpurun_out/c7_resstack python scripts/resstack_bench.py fused > gpurun_out/c7_ncu.log 2>&1
tail -2 gpurun_out/c7_ncu.log
