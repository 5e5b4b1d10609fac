set -x
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r01f_plain.json 2> gpurun_out/r01f_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_r01f_merger1e7.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r01f_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sweep_tile -s 3 -c 1 -f -o gpurun_out/prof_tile_r01f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r01f_ncu_full.log 2>&1
ls -la gpurun_out/prof_tile_r01f.ncu-rep
