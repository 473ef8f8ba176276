#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
timeout 300 python bench.py --steps 20 --warmup 3 --headline-only --no-cpu-full-batch > gpurun_out/c19b_bench_$i.json 2>/dev/null
python - gpurun_out/c19b_bench_$i.json <<PY
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(round(d["ms_per_step"],2), "e2e", round(d["e2e"]["ms_per_step"],2), d["extra"]["step_ms_spread"], d["clocks"]["sm_mhz"], d["clocks"]["sm_mhz_last_samples"])
PY
done
K='regex:conv_tc|conv_simt|resstack|attention|layernorm|embed|durations|length_regulate|variance_head|pack_|softmax_rows|rowbias|conv_post|add_positions|wav_to'
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$K" -s 500 -c 340 --csv --log-file gpurun_out/c19b_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --headline-only > gpurun_out/c19b_b.log 2>&1
wc -l gpurun_out/c19b_launches.csv
