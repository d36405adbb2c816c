#!/bin/bash
T=r3f
B="bench.py --no-cpu --no-extras --steps 20 --warmup 5"
timeout 300 python $B > gpurun_out/${T}_prio.log 2>&1
PNVO_GRAPH_PRIORITY=0 timeout 300 python $B > gpurun_out/${T}_noprio.log 2>&1
timeout 300 python $B > gpurun_out/${T}_prio2.log 2>&1
grep -o '"ms_per_step": [0-9.]*' gpurun_out/${T}_*.log
