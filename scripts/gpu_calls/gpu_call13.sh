#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/flip_census.py 3 2>&1 | grep -v Warn | tee gpurun_out/c13_flip_census.jsonl | cut -c1-700
