# full ncu capture of the tcgen05 conv kernel on representative HiFi-GAN layers (vocoder-only step); TAG names the outputs
# conv_tc launch order in a vocoder-only step: 0 conv_pre; 1-2 ups0; 3-20 stage0 resblocks; 21-22 ups1; 23-40 stage1 (23-28 k3, 29-34 k7, 35-40 k11); 41-42 ups2; 43-60 stage2; 61-62 ups3; 63-80 stage3
# (gpurun brings back at most 64 MiB: one launch per report, plus the raw / source pages as gzipped csv)
TAG=${TAG:-final}
cap() {  # name, skip
  ncu --set full --clock-control none --import-source on -k regex:conv_tc --profile-from-start off -s $2 -c 1 -o gpurun_out/prof_${TAG}_$1 -f python scripts/profile_step.py --vocoder-only > gpurun_out/ncu_$1.log 2>&1
  ncu -i gpurun_out/prof_${TAG}_$1.ncu-rep --page raw --csv 2>/dev/null | gzip > gpurun_out/prof_${TAG}_$1_raw.csv.gz
  ncu -i gpurun_out/prof_${TAG}_$1.ncu-rep --page source --csv --print-source cuda,sass 2>/dev/null | gzip > gpurun_out/prof_${TAG}_$1_source.csv.gz
}
cap s1k3 23
cap s1k3res 24
cap s1k11 35
cap s3k3 63
rm -f gpurun_out/prof_${TAG}_s1k3res.ncu-rep gpurun_out/prof_${TAG}_s1k11.ncu-rep   # keep two reports within the size limit
ls -la gpurun_out/
