#!/bin/bash
# Round-end evidence on one B200: GPU suite, smoke, launch list + ncu --set full of ONE training step, K4 stem kernels, bench lines.
T=${1:-r02f}
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/${T}_pytest.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/${T}_smoke.log 2>&1
PNVO_GRAPHS=0 PNVO_PROFILE_STEP=1 timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-extras --no-prefetch > gpurun_out/${T}_ncu.log 2>&1
PNVO_GRAPHS=0 PNVO_PROFILE_STEP=1 timeout 900 ncu --profile-from-start off --set full --clock-control none -f -o /tmp/${T}_step python bench.py --steps 1 --warmup 3 --no-cpu --no-extras --no-prefetch > gpurun_out/${T}_ncu_full.log 2>&1
python tools/ncu_summary.py /tmp/${T}_step.ncu-rep > gpurun_out/${T}_ncu_full_step.txt 2>&1
timeout 400 ncu --set full --clock-control none -k regex:conv_direct1 -c 2 -f -o /tmp/${T}_k4_stem python tools/k4_update.py 0 > gpurun_out/${T}_k4_ncu.log 2>&1
python tools/ncu_summary.py /tmp/${T}_k4_stem.ncu-rep > gpurun_out/${T}_ncu_full_k4_stem.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.log 2>&1
ls -la gpurun_out/${T}_*
tail -2 gpurun_out/${T}_pytest.log; tail -1 gpurun_out/${T}_smoke.log
