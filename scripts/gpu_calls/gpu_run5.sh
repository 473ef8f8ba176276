python scripts/tc_bringup.py > gpurun_out/tc_bringup12.log 2>&1
rm -f gpurun_out/parity_report.jsonl
python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest_gpu_tc12.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_tc12.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tc12.log 2> gpurun_out/bench_tc12.err
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_tc_v12.csv python scripts/profile_step.py > gpurun_out/ncu1.log 2>&1
cat gpurun_out/tc_bringup12.log | tail -14; tail -3 gpurun_out/pytest_gpu_tc12.log; grep hifigan gpurun_out/parity_report.jsonl; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_tc12.log') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['roofline']['achieved'], d['extra']['fastspeech2_only_ms_per_step'], d['e2e']['value'])
PY
tail -3 gpurun_out/bench_tc12.err
