set -x
python scripts/dbg_fast_pow2.py > gpurun_out/dbg_pow2.log 2>&1; cat gpurun_out/dbg_pow2.log
timeout 1200 python -m pytest tests/test_gpu_fast.py -q -m gpu > gpurun_out/t_fast5.log 2>&1; echo "rc=$?" >> gpurun_out/t_fast5.log
grep -E "FAILED|ERROR|passed|failed|rc=|^E  " gpurun_out/t_fast5.log | tail -12
python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b_i.json 2> gpurun_out/b_i.err; cut -c1-400 gpurun_out/b_i.json
