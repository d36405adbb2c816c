#!/bin/bash
T=r3b
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29611 tests/multigpu/peer_adam_check.py > gpurun_out/${T}_peer_check.log 2>&1
echo "peer check rc=$?"
tail -5 gpurun_out/${T}_peer_check.log
timeout 400 $TR --master-port 29612 bench.py --gpus 2 --no-cpu --steps 20 --warmup 5 > gpurun_out/${T}_bench2_peer.log 2>&1
echo "bench peer rc=$?"
PNVO_PEER_ADAM=0 timeout 400 $TR --master-port 29613 bench.py --gpus 2 --no-cpu --steps 20 --warmup 5 > gpurun_out/${T}_bench2_nccl.log 2>&1
echo "bench nccl rc=$?"
grep -o '"ms_per_step": [0-9.]*' gpurun_out/${T}_bench2_peer.log gpurun_out/${T}_bench2_nccl.log
timeout 200 python -m pytest tests -m gpu -q -x -k "state_dict_resumes or peer_memory" 2>&1 | tail -5
