#!/bin/bash
T=r3s
PNVO_GRAPHS=0 PNVO_PROFILE_STEP=1 timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_r50_launches.csv python bench.py --model r50_8ch --steps 1 --warmup 3 --no-cpu --no-extras --no-prefetch > gpurun_out/${T}_ncu.log 2>&1
tail -2 gpurun_out/${T}_ncu.log | cut -c1-300
