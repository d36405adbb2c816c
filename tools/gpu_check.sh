#!/bin/bash
# One GPU-box call that re-validates a change: the whole GPU suite, smoke(), a launch list of the last training steps and
# the default bench line.  Usage: gpurun --timeout 1500 -- 'bash tools/gpu_check.sh TAG'   (outputs under gpurun_out/TAG_*)
T=${1:-chk}
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/${T}_pytest.log
timeout 200 python __graft_entry__.py --smoke > gpurun_out/${T}_smoke.log 2>&1
PNVO_GRAPHS=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1300 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-prefetch > gpurun_out/${T}_ncu.log 2>&1
timeout 300 python bench.py ${BENCH_ARGS:---no-cpu} > gpurun_out/${T}_bench.log 2>&1
