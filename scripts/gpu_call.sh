set -x
timeout 2000 python -m pytest tests -q -m gpu > gpurun_out/t_all5.log 2>&1; echo "rc=$?" >> gpurun_out/t_all5.log
grep -E "FAILED|ERROR|passed|failed|rc=|^E  " gpurun_out/t_all5.log | tail -30
python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b_g.json 2> gpurun_out/b_g.err && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02g.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02g_ncu_list.log 2>&1
cat gpurun_out/b_g.json
timeout 900 python scripts/bench_configs.py r02 > gpurun_out/configs.log 2>&1; tail -6 gpurun_out/configs.log
