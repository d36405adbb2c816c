timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/t49_pytest.log
PNVO_GRAPHS=0 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"topdown|pack_w_multi|unpack_dw_multi" -s 6 -c 3 --csv --log-file gpurun_out/t49_k.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-prefetch > gpurun_out/t49_ncu.log 2>&1
for rep in 1 2; do timeout 200 python bench.py --no-cpu --steps 20 2>&1 | tail -1 | grep -o '"ms_per_step": [0-9.]*' | head -2 >> gpurun_out/t49_bench.log; done
