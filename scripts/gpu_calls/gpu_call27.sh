#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
( time timeout 380 python -m pytest tests -q -m gpu --durations=5 ) > gpurun_out/f1_tests.txt 2>&1
tail -12 gpurun_out/f1_tests.txt
