timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/b_p9.json 2> gpurun_out/b_p9.err; cut -c1-200 gpurun_out/b_p9.json
timeout 900 python -m pytest tests/test_gpu_fast.py -x -q -m gpu > gpurun_out/t_fast_p9.log 2>&1; tail -3 gpurun_out/t_fast_p9.log
