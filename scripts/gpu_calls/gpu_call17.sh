#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-full-batch > gpurun_out/c17_bench.json 2> gpurun_out/c17_bench.err
tail -c 300 gpurun_out/c17_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/c17_bench.json").read().strip().splitlines()[-1])
print(round(d["ms_per_step"],2), "ms/step e2e", round(d["e2e"]["ms_per_step"],2), "warmup", d["warmup"], "frac", round(d["roofline"]["frac"],3), d["clocks"])
for k,v in d["extra"]["configs"].items(): print(k, {a: (round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a!="workload"})
PY
K='regex:conv_tc|conv_simt|resstack|attention|layernorm|embed|durations|length_regulate|variance_head|pack_|softmax_rows|rowbias|conv_post|add_positions|wav_to'
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$K" -s 560 -c 260 --csv --log-file gpurun_out/c17_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --headline-only > gpurun_out/c17_b.log 2>&1
wc -l gpurun_out/c17_launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_fused -s 8 -c 1 -o gpurun_out/c17_attention_fused python bench.py --steps 1 --warmup 3 --no-cpu-baseline --headline-only > gpurun_out/c17_ncu_att.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 330 -c 1 -o gpurun_out/c17_conv_tc_s1k11 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --headline-only > gpurun_out/c17_ncu_conv.log 2>&1
ls -la gpurun_out/*.ncu-rep
