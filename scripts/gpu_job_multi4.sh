#!/bin/bash
T="timeout 300"
for N in 1 2 4; do
$T python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N scripts/audit_configs.py --config 5 > gpurun_out/audit5_n$N.json 2> gpurun_out/audit5_n$N.err
cut -c1-330 gpurun_out/audit5_n$N.json; grep -i "error\|Traceback" gpurun_out/audit5_n$N.err | head -3
done
$T python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err
cut -c1-200 gpurun_out/bench_n4.json
$T python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print(d['value'], d['e2e']['value'], d['e2e_track_u8'], d['roofline']['frac'], d['cpu_baseline'])"
