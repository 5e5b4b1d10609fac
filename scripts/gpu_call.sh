set -x
timeout 2000 python -m pytest tests -q -m gpu -v > gpurun_out/t_all3.log 2>&1; echo "rc=$?" >> gpurun_out/t_all3.log
grep -E "PASSED|FAILED|ERROR|SKIPPED|passed|failed|rc=" gpurun_out/t_all3.log | tail -70
