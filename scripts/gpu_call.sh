for s in 1 7; do
timeout 200 python scripts/fuzz_fast.py 35 400003 $s > gpurun_out/ff_new_$s.log 2>&1
TOYGPU_LIB=$PWD/toycluster_b200/variants/libtoygpu_a_fp64sum.so timeout 200 python scripts/fuzz_fast.py 35 400003 $s > gpurun_out/ff_old_$s.log 2>&1
done
grep -h "within\|shift=1" gpurun_out/ff_new_1.log gpurun_out/ff_old_1.log gpurun_out/ff_new_7.log gpurun_out/ff_old_7.log | cut -c1-175
