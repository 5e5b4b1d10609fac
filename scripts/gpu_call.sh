set -x
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/t_all7.log 2>&1; echo "rc=$?" >> gpurun_out/t_all7.log
grep -E "FAILED|ERROR|passed|failed|rc=|^E  " gpurun_out/t_all7.log | tail -20
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/b_fast2.json 2> gpurun_out/b_fast2.err; cut -c1-330 gpurun_out/b_fast2.json; tail -2 gpurun_out/b_fast2.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/b_final1b.json 2> gpurun_out/b_final1b.err; cut -c1-330 gpurun_out/b_final1b.json
