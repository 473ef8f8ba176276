# end-of-iteration validation: parity tests, smoke, bench (both arms), launch list
rm -f gpurun_out/parity_report.jsonl
python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest_gpu_final.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_final.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke_final.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_final.log 2> gpurun_out/bench_final.err
echo "bench exit $?" >> gpurun_out/bench_final.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_final.log 2> gpurun_out/bench_ref_final.err
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_final.csv python scripts/profile_step.py > gpurun_out/ncu1.log 2>&1
tail -3 gpurun_out/pytest_gpu_final.log; tail -2 gpurun_out/smoke_final.log; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_final.log') if l.startswith('{')][-1])
print("value", d['value'], "ms", d['ms_per_step'], "conv TF/s", d['roofline']['achieved'], "frac", d['roofline']['frac'], "fs2 ms", d['extra']['fastspeech2_only_ms_per_step'], "mel fps", d['extra']['fastspeech2_only_mel_frames_per_s'], "e2e", d['e2e']['value'])
print("cpu", d['cpu_baseline']); print("clocks", d['clocks'])
r=json.loads([l for l in open('gpurun_out/bench_ref_final.log') if l.startswith('{')][-1]); print("ref arm", r['value'], r['cpu_baseline']['cores'], r['ms_per_step'])
PY
tail -2 gpurun_out/bench_final.err
