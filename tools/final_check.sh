timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/t56_pytest.log
timeout 300 python bench.py --no-cpu > gpurun_out/t56_bench.log 2>&1
timeout 300 python bench.py --no-cpu --depth fp32 > gpurun_out/t56_bench_fp32.log 2>&1
