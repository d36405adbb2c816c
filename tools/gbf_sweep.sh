for kb in 48 75 100; do
PNVO_GBF_SMEM_KB=$kb PNVO_GRAPHS=0 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gn_bwd_fused -s 40 -c 20 --csv --log-file gpurun_out/t44_gbf_$kb.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-prefetch > gpurun_out/t44_ncu_$kb.log 2>&1
done
for kb in 48 75 100; do PNVO_GBF_SMEM_KB=$kb timeout 200 python bench.py --no-cpu 2>&1 | tail -1 | cut -c1-220 > gpurun_out/t44_bench_$kb.log; done
