#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/c11_bench.json 2> gpurun_out/c11_bench.err
tail -c 600 gpurun_out/c11_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/c11_bench.json").read().strip().splitlines()[-1])
print(round(d["ms_per_step"],2), "ms/step e2e", round(d["e2e"]["ms_per_step"],2), "frac", round(d["roofline"]["frac"],3), d["clocks"])
for k,v in d["extra"]["configs"].items(): print(k, {a: (round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a!="workload"})
print(d["cpu_baseline"])
PY
K='regex:conv_tc|conv_simt|resstack|attention|layernorm|embed|durations|length_regulate|variance_head|pack_|softmax_rows|rowbias|conv_post|add_positions|wav_to'
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$K" -s 500 -c 300 --csv --log-file gpurun_out/c11_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --headline-only > gpurun_out/c11_b.log 2>&1
wc -l gpurun_out/c11_launches.csv
