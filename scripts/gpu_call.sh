set -x
timeout 900 python -m pytest tests -q -m gpu -k "blocks" > gpurun_out/t_blocks.log 2>&1; echo "rc=$?" >> gpurun_out/t_blocks.log
grep -E "FAILED|ERROR|passed|failed|rc=|^E  " gpurun_out/t_blocks.log | tail -20
