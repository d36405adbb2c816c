#!/bin/bash
T=r3k
PNVO_GRAPHS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_k4_launches.csv python tools/k4_update.py 1 > gpurun_out/${T}_k4.log 2>&1
tail -2 gpurun_out/${T}_k4.log
timeout 300 python tools/k4_update.py 3 2>&1 | tail -1
