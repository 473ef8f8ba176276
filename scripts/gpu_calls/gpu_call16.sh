#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "resstack" -x ) > gpurun_out/c16_tests.txt 2>&1
tail -8 gpurun_out/c16_tests.txt
timeout 200 python scripts/pair_bench.py 2>&1 | tee gpurun_out/c16_pair_bench.txt
for pm in 4 12 ; do for km in 3 7 11; do
timeout 200 python bench.py --steps 10 --warmup 3 --headline-only --voc-pair-mask $pm --voc-pair-kmax $km --voc-fused-mask $([ $pm = 12 ] && echo 0 || echo 8) 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pair_mask $pm kmax $km', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), d['gpu_launches'])"
done; done
