#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv --log-file gpurun_out/c5_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/c5_b.log 2>&1
tail -2 gpurun_out/c5_b.log | cut -c1-300
wc -l gpurun_out/c5_launches.csv
