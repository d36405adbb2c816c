#!/bin/bash
T=r3j
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29611 tests/multigpu/peer_adam_check.py > gpurun_out/${T}_peer_check.log 2>&1
echo "peer check rc=$?"; tail -1 gpurun_out/${T}_peer_check.log
timeout 400 $TR --master-port 29612 bench.py --gpus 8 --no-cpu --steps 20 --warmup 5 > gpurun_out/${T}_r18_8gpu.log 2>&1
timeout 400 $TR --master-port 29613 bench.py --gpus 8 --no-cpu --steps 10 --warmup 3 --model r50_8ch > gpurun_out/${T}_r50_8gpu.log 2>&1
grep -o '"ms_per_step": [0-9.]*' gpurun_out/${T}_*.log
grep -o '"grad_exchange": {[^}]*}' gpurun_out/${T}_*.log
