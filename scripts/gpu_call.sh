cp toycluster_b200/csrc/tile_fast.cuh gpurun_out/r02s_tile_fast.cuh
cuobjdump -xelf all toycluster_b200/libtoygpu.so > /dev/null 2>&1; mv toygpu.sm_100a.cubin gpurun_out/r02s.cubin 2>/dev/null
timeout 900 bash scripts/prof_round.sh r02s 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 --full-relaxation > gpurun_out/b_final3_1gpu.json 2> gpurun_out/b_final3_1gpu.err; cut -c1-200 gpurun_out/b_final3_1gpu.json
