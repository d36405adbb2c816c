#!/bin/bash
T=r3a
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/${T}_pytest.log
timeout 300 python bench.py --no-cpu --steps 20 --warmup 5 > gpurun_out/${T}_bench.log 2>&1
PNVO_RASTER_CONCAT=0 timeout 300 python bench.py --no-cpu --steps 20 --warmup 5 > gpurun_out/${T}_bench_noconcat.log 2>&1
PNVO_GRAPHS=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1300 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-prefetch > gpurun_out/${T}_ncu.log 2>&1
tail -3 gpurun_out/${T}_pytest.log
grep -o '"ms_per_step": [0-9.]*' gpurun_out/${T}_bench.log gpurun_out/${T}_bench_noconcat.log
