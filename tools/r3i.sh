#!/bin/bash
T=r3i
B="bench.py --no-cpu --no-extras --steps 20 --warmup 5"
timeout 300 python $B > gpurun_out/${T}_fused.log 2>&1
PNVO_FUSED_STATS=0 timeout 300 python $B > gpurun_out/${T}_sep.log 2>&1
timeout 300 python $B > gpurun_out/${T}_fused2.log 2>&1
PNVO_FUSED_STATS=0 timeout 300 python $B > gpurun_out/${T}_sep2.log 2>&1
grep -o '"ms_per_step": [0-9.]*' gpurun_out/${T}_*.log
