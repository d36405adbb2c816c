#!/bin/bash
T=r3q
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/${T}_pytest.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/${T}_smoke.log 2>&1
PNVO_GRAPHS=0 PNVO_PROFILE_STEP=1 timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-extras --no-prefetch > gpurun_out/${T}_ncu.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.log 2>&1
tail -2 gpurun_out/${T}_pytest.log; tail -1 gpurun_out/${T}_smoke.log
