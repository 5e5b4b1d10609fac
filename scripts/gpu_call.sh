set -x
timeout 900 python -m pytest tests/test_gpu_fast.py -x -q -k "not full_size" > gpurun_out/t_fast.log 2>&1; echo "rc=$?" >> gpurun_out/t_fast.log
tail -15 gpurun_out/t_fast.log
python scripts/time_variants.py > gpurun_out/variants.log 2>&1; cat gpurun_out/variants.log
bash scripts/prof_round.sh r02b
