#!/bin/bash
# usage: gpu_job_multi.sh N   — bench + long-video audit (config 5) on N GPUs of one box
N=$1
T="timeout 300"
$T python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
cut -c1-200 gpurun_out/bench_n$N.json; tail -n 3 gpurun_out/bench_n$N.err
$T python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/audit_configs.py --config 5 > gpurun_out/audit5_n$N.json 2> gpurun_out/audit5_n$N.err
cat gpurun_out/audit5_n$N.json | cut -c1-400; tail -n 3 gpurun_out/audit5_n$N.err
