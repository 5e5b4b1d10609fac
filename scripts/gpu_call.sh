set -x
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/t_all7.log 2>&1; echo "rc=$?" >> gpurun_out/t_all7.log
grep -E "FAILED|ERROR|passed|failed|rc=|^E  " gpurun_out/t_all7.log | tail -20
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/b_final1b.json 2> gpurun_out/b_final1b.err; cut -c1-330 gpurun_out/b_final1b.json
python scripts/time_variants.py > gpurun_out/variants3.log 2>&1; cut -c1-120 gpurun_out/variants3.log
