#!/bin/bash
T=r3n
PNVO_GRAPHS=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-extras --no-prefetch > gpurun_out/${T}_ncu.log 2>&1
