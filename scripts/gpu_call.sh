# scratch entry point for `gpurun -- 'bash scripts/gpu_call.sh'` (edited per call during development)
set -x
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/t_all.log 2>&1; echo "rc=$?" >> gpurun_out/t_all.log
grep -E "FAILED|ERROR|passed|failed|rc=|^E  " gpurun_out/t_all.log | tail -10
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/b.json 2> gpurun_out/b.err; cut -c1-200 gpurun_out/b.json
