#!/bin/bash
T=r3g
B="bench.py --no-cpu --no-extras --steps 20 --warmup 5"
for k in 0 1 3 7 4 6; do
PNVO_DIAG_SKIP_INPUT=$k timeout 300 python $B > gpurun_out/${T}_skip$k.log 2>&1
done
PNVO_DIAG_SKIP_INPUT=7 timeout 300 python $B --no-prefetch > gpurun_out/${T}_skip7_noprefetch.log 2>&1
grep -o '"ms_per_step": [0-9.]*' gpurun_out/${T}_*.log
