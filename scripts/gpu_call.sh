python scripts/dbg_fast_pow2.py > gpurun_out/dbg_pow2.log 2>&1; tail -9 gpurun_out/dbg_pow2.log | cut -c1-900
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b_j.json 2> gpurun_out/b_j.err; cut -c1-330 gpurun_out/b_j.json
