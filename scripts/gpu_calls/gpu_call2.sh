#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
( time python -m pytest tests -x -q -m gpu --durations=8 ) > gpurun_out/c2_tests.txt 2>&1
tail -30 gpurun_out/c2_tests.txt
