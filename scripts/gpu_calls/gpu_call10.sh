#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -q -m gpu --deselect tests/test_gpu_model.py::test_fastspeech2_full_size_vs_oracle --deselect tests/test_gpu_model.py::test_fastspeech2_config4_shard_vs_oracle --durations=6 ) > gpurun_out/c10_tests.txt 2>&1
tail -14 gpurun_out/c10_tests.txt
