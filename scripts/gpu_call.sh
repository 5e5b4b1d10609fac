set -x
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/t_all12.log 2>&1; echo "rc=$?" >> gpurun_out/t_all12.log
grep -E "FAILED|ERROR|passed|failed|rc=|^E  " gpurun_out/t_all12.log | tail -10
timeout 900 python bench.py --steps 20 --warmup 5 --full-relaxation > gpurun_out/b_final5_1gpu.json 2> gpurun_out/b_final5_1gpu.err; cut -c1-200 gpurun_out/b_final5_1gpu.json
