ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_tc_v1.csv python scripts/profile_step.py > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc --profile-from-start off -s 35 -c 2 -o gpurun_out/prof_tc_s1k11 -f python scripts/profile_step.py --vocoder-only > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc --profile-from-start off -s 75 -c 2 -o gpurun_out/prof_tc_s3k11 -f python scripts/profile_step.py --vocoder-only > gpurun_out/ncu3.log 2>&1
tail -2 gpurun_out/ncu1.log gpurun_out/ncu2.log gpurun_out/ncu3.log; wc -l gpurun_out/launches_tc_v1.csv; ls -la gpurun_out
