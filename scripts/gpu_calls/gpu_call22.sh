#!/bin/bash
mkdir -p gpurun_out
python scripts/pair_bench.py 2>&1 | tee gpurun_out/c22_pair_bench.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:resstack -s 1 -c 1 -o gpurun_out/c22_pair_k3 python scripts/pair_bench.py one > gpurun_out/c22_ncu.log 2>&1
ls -la gpurun_out/c22*.ncu-rep
