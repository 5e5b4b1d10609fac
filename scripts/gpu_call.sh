timeout 600 python scripts/time_variants.py > gpurun_out/variants14.log 2>&1; cat gpurun_out/variants14.log | cut -c1-100
