set -x
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu > gpurun_out/t_multi3.log 2>&1; echo "rc=$?" >> gpurun_out/t_multi3.log
grep -E "FAILED|ERROR|passed|failed|rc=|^E  " gpurun_out/t_multi3.log | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/b_final4_2gpu.json 2> gpurun_out/b_final4_2gpu.err; cut -c1-200 gpurun_out/b_final4_2gpu.json
