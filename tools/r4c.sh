#!/bin/bash
T=r4c
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > gpurun_out/${T}_pytest.log; tail -2 gpurun_out/${T}_pytest.log
B="bench.py --no-cpu --no-extras --steps 20 --warmup 5"
timeout 300 python $B > gpurun_out/${T}_base.log 2>&1
PNVO_FUSED_DY_SUMS=0 timeout 300 python $B > gpurun_out/${T}_sep.log 2>&1
grep -o '"ms_per_step": [0-9.]*' gpurun_out/${T}_*.log
