# full ncu capture of the tcgen05 conv kernel on representative HiFi-GAN layers (vocoder-only step); TAG names the outputs
# conv_tc launch order in a vocoder-only step: 0 conv_pre; 1-2 ups0; 3-20 stage0 resblocks; 21-22 ups1; 23-40 stage1 (23-28 k3, 29-34 k7, 35-40 k11); 41-42 ups2; 43-60 stage2; 61-62 ups3; 63-80 stage3
TAG=${TAG:-final}
ncu --set full --clock-control none --import-source on -k regex:conv_tc --profile-from-start off -s 23 -c 2 -o gpurun_out/prof_${TAG}_s1k3 -f python scripts/profile_step.py --vocoder-only > gpurun_out/ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc --profile-from-start off -s 35 -c 1 -o gpurun_out/prof_${TAG}_s1k11 -f python scripts/profile_step.py --vocoder-only > gpurun_out/ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc --profile-from-start off -s 63 -c 2 -o gpurun_out/prof_${TAG}_s3k3 -f python scripts/profile_step.py --vocoder-only > gpurun_out/ncu_c.log 2>&1
ls -la gpurun_out/*.ncu-rep
