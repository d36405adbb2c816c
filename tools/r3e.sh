#!/bin/bash
T=r3e
timeout 300 python bench.py --no-cpu --no-extras --steps 20 --warmup 5 --no-prefetch > gpurun_out/${T}_noprefetch.log 2>&1
timeout 300 python bench.py --no-cpu --no-extras --steps 20 --warmup 5 > gpurun_out/${T}_prefetch.log 2>&1
grep -o '"ms_per_step": [0-9.]*' gpurun_out/${T}_*.log
