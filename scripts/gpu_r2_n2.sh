#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 2> gpurun_out/r02_bench_n2.err | tail -n 1 > gpurun_out/r02_bench_n2.json; cut -c1-1200 gpurun_out/r02_bench_n2.json; tail -3 gpurun_out/r02_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>/dev/null | tail -n 1 | cut -c1-600
