#!/bin/bash
mkdir -p gpurun_out
python scripts/conv_layer_bench.py 2>&1 | tee gpurun_out/c20_conv_layer_bench.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 1 -c 1 -o gpurun_out/c20_s1k3 python scripts/conv_layer_bench.py one > gpurun_out/c20_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 5 -c 1 -o gpurun_out/c20_s1k7 python scripts/conv_layer_bench.py one > gpurun_out/c20_ncu2.log 2>&1
ls -la gpurun_out/c20*.ncu-rep
