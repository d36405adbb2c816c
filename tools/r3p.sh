#!/bin/bash
T=r3p
B="bench.py --no-cpu --no-extras --steps 20 --warmup 5"
PNVO_PREFETCH_AT=bwd timeout 300 python $B > gpurun_out/${T}_bwd.log 2>&1
timeout 300 python $B > gpurun_out/${T}_start.log 2>&1
grep -o '"ms_per_step": [0-9.]*' gpurun_out/${T}_*.log
