timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu > gpurun_out/t_multi4.log 2>&1; echo "rc=$?" >> gpurun_out/t_multi4.log
grep -E "FAILED|ERROR|passed|failed|rc=|^E  " gpurun_out/t_multi4.log | tail -5
