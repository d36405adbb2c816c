timeout 400 python -m pytest tests -m gpu -x -q -k "groupnorm or golden or fused_train or block_by_block or reproducible" 2>&1 | tail -5 > gpurun_out/t47_pytest.log
for v in 0 1; do
PNVO_POOL_BWD_2X2=$v PNVO_GRAPHS=0 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pool_bwd -s 2 -c 2 --csv --log-file gpurun_out/t47_pool_$v.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-prefetch > gpurun_out/t47_ncu.log 2>&1
done
for rep in 1 2; do for v in 0 1; do
PNVO_POOL_BWD_2X2=$v timeout 200 python bench.py --no-cpu --steps 20 2>&1 | tail -1 | grep -o '"ms_per_step": [0-9.]*' | head -1 >> gpurun_out/t47_bench_$v.log; done; done
