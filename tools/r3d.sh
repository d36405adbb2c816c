#!/bin/bash
T=r3d
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
B="bench.py --gpus 2 --no-cpu --steps 20 --warmup 5"
timeout 300 $TR --master-port 29611 tests/multigpu/peer_adam_check.py > gpurun_out/${T}_peer_check.log 2>&1
echo "peer check rc=$?"; tail -2 gpurun_out/${T}_peer_check.log
timeout 300 $TR --master-port 29612 $B > gpurun_out/${T}_peer.log 2>&1
PNVO_PEER_STATS=0 PNVO_PEER_ADAM=0 timeout 300 $TR --master-port 29613 $B > gpurun_out/${T}_nccl.log 2>&1
grep -o '"ms_per_step": [0-9.]*' gpurun_out/${T}_*.log
