set -x
timeout 2000 python -m pytest tests -q -m gpu > gpurun_out/t_all6.log 2>&1; echo "rc=$?" >> gpurun_out/t_all6.log
grep -E "FAILED|ERROR|passed|failed|rc=|^E  " gpurun_out/t_all6.log | tail -30
bash scripts/prof_round.sh r02h
cat gpurun_out/r02h_plain.json
