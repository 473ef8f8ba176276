#!/bin/bash
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv_post -s 2 -c 1 -o gpurun_out/c28_conv_post -f python scripts/conv_post_one.py > gpurun_out/c28_ncu.log 2>&1
tail -2 gpurun_out/c28_ncu.log; ls -la gpurun_out/c28_conv_post.ncu-rep
