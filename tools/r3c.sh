#!/bin/bash
T=r3c
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
B="bench.py --gpus 2 --no-cpu --steps 20 --warmup 5"
timeout 300 $TR --master-port 29612 $B > gpurun_out/${T}_peer.log 2>&1
PNVO_DIAG_LOCAL_STATS=1 timeout 300 $TR --master-port 29613 $B > gpurun_out/${T}_localstats.log 2>&1
PNVO_DIAG_LOCAL_STATS=1 PNVO_DIAG_NO_EXCHANGE=1 timeout 300 $TR --master-port 29614 $B > gpurun_out/${T}_noexch.log 2>&1
CUDA_VISIBLE_DEVICES=1 timeout 300 python bench.py --no-cpu --no-extras --steps 20 --warmup 5 > gpurun_out/${T}_gpu1_alone.log 2>&1
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --no-cpu --no-extras --steps 20 --warmup 5 > gpurun_out/${T}_gpu0_alone.log 2>&1
grep -o '"ms_per_step": [0-9.]*' gpurun_out/${T}_*.log
timeout 200 python -m pytest tests -m gpu -q -x -k "peer_memory" > gpurun_out/${T}_pytest.log 2>&1
tail -3 gpurun_out/${T}_pytest.log
