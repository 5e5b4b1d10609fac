set -x
timeout 2000 python -m pytest tests/test_gpu_fast.py tests/test_gpu_golden.py tests/test_displaced_nodes.py tests/test_gpu_parity.py -q -m gpu -x > gpurun_out/t_sel2.log 2>&1; echo "rc=$?" >> gpurun_out/t_sel2.log
grep -E "FAILED|ERROR|passed|failed|rc=|^E  " gpurun_out/t_sel2.log | tail -30
bash scripts/prof_round.sh r02f
cat gpurun_out/r02f_plain.json
