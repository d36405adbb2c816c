for mb in 6 8; do
PNVO_GN_MINB=$mb PNVO_GRAPHS=0 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gn_apply -s 64 -c 32 --csv --log-file gpurun_out/t53_gn_$mb.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-prefetch > gpurun_out/t53_ncu.log 2>&1
done
for rep in 1 2; do for mb in 6 8; do PNVO_GN_MINB=$mb timeout 200 python bench.py --no-cpu --steps 20 2>&1 | tail -1 | grep -o '"ms_per_step": [0-9.]*' | head -1 >> gpurun_out/t53_bench_$mb.log; done; done
timeout 300 python -m pytest tests -m gpu -x -q -k "groupnorm or golden or split" 2>&1 | tail -3 > gpurun_out/t53_pytest.log
