#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 2> gpurun_out/bench_n2.err | tail -n 1 > gpurun_out/bench_n2.json
cut -c1-400 gpurun_out/bench_n2.json; tail -n 4 gpurun_out/bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2> gpurun_out/ref_n2.err | tail -n 1 | cut -c1-600
