set -x
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/t_all9.log 2>&1; echo "rc=$?" >> gpurun_out/t_all9.log
grep -E "FAILED|ERROR|passed|failed|rc=|^E  " gpurun_out/t_all9.log | tail -20
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/b_p3.json 2> gpurun_out/b_p3.err; cut -c1-200 gpurun_out/b_p3.json
TOYGPU_NO_SIDE_STREAM=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/b_p3_noside.json 2> gpurun_out/b_p3_noside.err; cut -c1-200 gpurun_out/b_p3_noside.json
