ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_r02u.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02u_ncu_list.log 2>&1
timeout 600 python scripts/bench_configs.py r02d > gpurun_out/configs_r02d.log 2>&1; tail -1 gpurun_out/configs_r02d.log | cut -c1-200
