for cfg in "0 1" "0 4" "1 4" "1 1"; do set -- $cfg
PNVO_GN_EARLY=$1 PNVO_GN_MIN_ITEMS=$2 PNVO_GRAPHS=0 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gn_apply -s 64 -c 32 --csv --log-file gpurun_out/t46_gn_$1_$2.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-prefetch > gpurun_out/t46_ncu.log 2>&1
done
for rep in 1 2; do for cfg in "0 1" "0 4" "1 4"; do set -- $cfg
PNVO_GN_EARLY=$1 PNVO_GN_MIN_ITEMS=$2 timeout 200 python bench.py --no-cpu --steps 20 2>&1 | tail -1 | grep -o '"ms_per_step": [0-9.]*' | head -1 >> gpurun_out/t46_bench_$1_$2.log; done; done
