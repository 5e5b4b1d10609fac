set -x
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 10 --warmup 3 --full-relaxation > gpurun_out/b_fast8.json 2> gpurun_out/b_fast8.err; cat gpurun_out/b_fast8.json; tail -3 gpurun_out/b_fast8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 4 --steps 10 --warmup 3 --no-e2e > gpurun_out/b_fast4.json 2> gpurun_out/b_fast4.err; cat gpurun_out/b_fast4.json
