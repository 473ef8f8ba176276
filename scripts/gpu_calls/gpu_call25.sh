#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_frontend.py -q -m gpu -x -k "conv_post or hifigan or synthesize or smoke or frontend or device_batches" ) > gpurun_out/c25_tests.txt 2>&1
tail -4 gpurun_out/c25_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python - <<PY
import torch, sys
sys.path.insert(0, ".")
from fastspeech2_b200 import ops
x = torch.randn(16, 259072, 32, device="cuda"); w = torch.randn(7, 32, device="cuda") * 0.1; b = torch.zeros(1, device="cuda")
for _ in range(3): ops.conv_post(x, w, b, 0.01)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): ops.conv_post(x, w, b, 0.01)
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 20 * 1e3
print(f"conv_post 16 x 259072 x 32: {us:.1f} us = {x.numel() * 4 / us / 1e6:.2f} TB/s of input read")
PY
timeout 300 python bench.py --steps 20 --warmup 3 --headline-only --no-cpu-full-batch > gpurun_out/c25_bench.json 2>/dev/null
python - <<PY
import json
d=json.loads(open("gpurun_out/c25_bench.json").read().strip().splitlines()[-1])
print(round(d["ms_per_step"],2), "e2e", round(d["e2e"]["ms_per_step"],2), d["extra"]["step_ms_spread"]["device_timed"], d["clocks"]["sm_mhz"], d["roofline"]["other_classes_ms"])
PY
