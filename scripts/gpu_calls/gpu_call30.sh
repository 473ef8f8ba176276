#!/bin/bash
mkdir -p gpurun_out
( timeout 100 python -m pytest tests/test_gpu_model.py tests/test_gpu_ops.py -q -m gpu -x -k "hifigan_vs_oracle or resstack or hifigan_golden" ) 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
