timeout 600 python scripts/time_variants.py > gpurun_out/variants15.log 2>&1; cat gpurun_out/variants15.log | cut -c1-100
timeout 600 python -m pytest tests/test_gpu_fast.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/t_p10.log 2>&1; tail -2 gpurun_out/t_p10.log
