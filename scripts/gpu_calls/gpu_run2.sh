python scripts/tc_bias_probe.py > gpurun_out/tc_bias.log 2>&1
python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest_gpu_tc.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_tc.log
cat gpurun_out/tc_bias.log; tail -40 gpurun_out/pytest_gpu_tc.log
