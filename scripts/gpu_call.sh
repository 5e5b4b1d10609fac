set -x
python scripts/time_variants.py > gpurun_out/variants2.log 2>&1; cat gpurun_out/variants2.log
