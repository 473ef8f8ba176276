#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
( timeout 400 python -m pytest tests -q -m gpu -k "batch_mode_front_end or real_checkpoint or device_batches or conv_post or synthesize_path" ) > gpurun_out/c26_tests.txt 2>&1
tail -5 gpurun_out/c26_tests.txt
cat gpurun_out/parity_report.jsonl | cut -c1-300
