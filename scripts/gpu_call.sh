timeout 600 python scripts/time_variants.py > gpurun_out/variants9.log 2>&1; cat gpurun_out/variants9.log | cut -c1-100
TOYGPU_VARIANTS=w4b7 TOYGPU_BENCH_WORKLOAD=merger_sub_1e7 timeout 600 python scripts/time_variants.py > gpurun_out/variants9s.log 2>&1; cat gpurun_out/variants9s.log | cut -c1-250
