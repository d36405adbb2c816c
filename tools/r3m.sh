#!/bin/bash
T=r3m
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/${T}_pytest.log
tail -4 gpurun_out/${T}_pytest.log
B="bench.py --no-cpu --no-extras --steps 20 --warmup 5"
timeout 300 python $B > gpurun_out/${T}_bench.log 2>&1
PNVO_S2_CLASSES=0 timeout 300 python $B > gpurun_out/${T}_bench_up.log 2>&1
grep -o '"ms_per_step": [0-9.]*' gpurun_out/${T}_*.log
timeout 300 python tools/k4_update.py 3 2>&1 | tail -1
