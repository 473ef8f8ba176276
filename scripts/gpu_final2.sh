#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/f2_bench_n2.json 2> gpurun_out/f2_bench_n2.err
tail -c 800 gpurun_out/f2_bench_n2.err
python - <<PY
import json
d=json.loads(open("gpurun_out/f2_bench_n2.json").read().strip().splitlines()[-1])
print(d["n_gpus"], round(d["ms_per_step"],2), "ms/step e2e", round(d["e2e"]["ms_per_step"],2), "value", d["value"], d["clocks"])
for k,v in d["extra"]["configs"].items(): print(k, {a: (round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a!="workload"})
PY
