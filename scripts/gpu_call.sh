set -x
nvidia-smi -L
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/t_all2.log 2>&1; echo "rc=$?" >> gpurun_out/t_all2.log
tail -40 gpurun_out/t_all2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/b_fast2.json 2> gpurun_out/b_fast2.err; cat gpurun_out/b_fast2.json; tail -5 gpurun_out/b_fast2.err
