#!/bin/bash
T=r3o
timeout 600 python tools/stem_wlo_probe.py > gpurun_out/${T}_probe.log 2>&1
cat gpurun_out/${T}_probe.log | grep STEM_WLO
B="bench.py --no-cpu --no-extras --steps 20 --warmup 5"
PNVO_STEM_WLO=0 timeout 300 python $B > gpurun_out/${T}_bench_nowlo.log 2>&1
timeout 300 python $B > gpurun_out/${T}_bench.log 2>&1
grep -o '"ms_per_step": [0-9.]*' gpurun_out/${T}_*.log
