timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/t54_pytest.log
for i in 1 2 3; do timeout 200 python -m pytest tests -m gpu -q -k "prefetched" 2>&1 | tail -1 >> gpurun_out/t54_prefetch.log; done
timeout 200 python __graft_entry__.py --smoke > gpurun_out/t54_smoke.log 2>&1
PNVO_GRAPHS=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 270 -c 160 --csv --log-file gpurun_out/t54_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-prefetch > gpurun_out/t54_ncu.log 2>&1
timeout 300 python bench.py > gpurun_out/t54_bench.log 2>&1
timeout 300 python bench.py --no-cpu --forward-only > gpurun_out/t54_bench_fwd.log 2>&1
