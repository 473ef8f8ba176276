#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 300 python -m pytest tests/test_gpu_model.py -x -q -k "tensor_core_encoder or frame_level or free_running_lj" 2>&1 | tail -12 > gpurun_out/c19_tests.txt
cat gpurun_out/c19_tests.txt
if grep -q "passed" gpurun_out/c19_tests.txt && ! grep -q "failed" gpurun_out/c19_tests.txt; then
  timeout 900 python scripts/flip_census.py 3 2>&1 | grep -v Warn | tee gpurun_out/c19_flip_census.jsonl | cut -c1-600
  python - <<PY
import sys, os, tempfile, torch, time
sys.path.insert(0, "/root/repo")
from fastspeech2_b200 import configs, synth, _lib as L
from fastspeech2_b200.model import FastSpeech2
pc, mc = configs.make_configs("LJSpeech", tempfile.mkdtemp())
sd = synth.fastspeech2_state_dict(pc, mc, seed=0)
for name, extra in (("fp32_cuda_cores", 0), ("tc_segmented", L.TC_ENCODER | L.TC_PREDICTORS)):
    m = FastSpeech2(pc, mc); m.load_state_dict(sd); m.tc_mask |= extra; m = m.to("cuda").eval()
    spk, texts, lens, Lm = synth.make_batch(16, 128, seed=0)
    a = [t.to("cuda") for t in (spk, texts, lens)]
    for _ in range(5): m(*a, Lm)
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): m(*a, Lm)
    e1.record(); torch.cuda.synchronize()
    print(name, "FastSpeech2-only ms/step", round(e0.elapsed_time(e1) / 20, 3))
PY
fi
