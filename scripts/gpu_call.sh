timeout 600 python scripts/time_variants.py > gpurun_out/variants12.log 2>&1; cat gpurun_out/variants12.log | cut -c1-200
timeout 900 python -m pytest tests/test_gpu_fast.py -x -q -m gpu > gpurun_out/t_fast_p4.log 2>&1; tail -3 gpurun_out/t_fast_p4.log
