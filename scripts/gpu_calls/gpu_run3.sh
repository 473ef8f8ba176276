rm -f gpurun_out/parity_report.jsonl
python -m pytest tests/test_gpu_model.py -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest_gpu_tc.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_tc.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tc.log 2> gpurun_out/bench_tc.err
tail -3 gpurun_out/pytest_gpu_tc.log; cat gpurun_out/parity_report.jsonl; tail -c 2500 gpurun_out/bench_tc.log; tail -3 gpurun_out/bench_tc.err
