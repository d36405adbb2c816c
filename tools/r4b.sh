#!/bin/bash
T=r4b
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > gpurun_out/${T}_pytest.log; tail -2 gpurun_out/${T}_pytest.log
B="bench.py --no-cpu --no-extras --steps 20 --warmup 5"
timeout 300 python $B > gpurun_out/${T}_base.log 2>&1
grep -o '"ms_per_step": [0-9.]*' gpurun_out/${T}_*.log
PNVO_GRAPHS=0 PNVO_PROFILE_STEP=1 timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-extras --no-prefetch > gpurun_out/${T}_ncu.log 2>&1
