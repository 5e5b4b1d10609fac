set -x
python scripts/sanitize_small.py > gpurun_out/san_plain.log 2>&1 && timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_small.py > gpurun_out/san_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/san_memcheck.log; tail -6 gpurun_out/san_memcheck.log
python scripts/handbacks.py merger_sub_1e7 > gpurun_out/handbacks_sub.log 2>&1; cat gpurun_out/handbacks_sub.log
timeout 1500 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/b_ref.json 2> gpurun_out/b_ref.err; cat gpurun_out/b_ref.json; tail -3 gpurun_out/b_ref.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/b_final1.json 2> gpurun_out/b_final1.err; cat gpurun_out/b_final1.json
