#!/bin/bash
T=r4a
B="bench.py --no-cpu --no-extras --steps 20 --warmup 5"
PNVO_DIAG_SKIP_WGRAD=1 timeout 300 python $B > gpurun_out/${T}_nowgrad.log 2>&1
PNVO_SIDE_LANE=0 timeout 300 python $B > gpurun_out/${T}_noside.log 2>&1
timeout 300 python $B > gpurun_out/${T}_base.log 2>&1
grep -o '"ms_per_step": [0-9.]*' gpurun_out/${T}_*.log
