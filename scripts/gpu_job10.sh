#!/bin/bash
T="timeout 200"
$T python -m pytest tests -m gpu -x -q 2>&1 | tail -4
$T python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_now.json 2> gpurun_out/bench_now.err; cut -c1-250 gpurun_out/bench_now.json
$T ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:logmel --csv --log-file gpurun_out/logmel_launches.csv python scripts/audit_configs.py --config 3 --batches 64,512 > gpurun_out/ncu.log 2>&1
grep -c logmel gpurun_out/logmel_launches.csv; tail -n 4 gpurun_out/logmel_launches.csv | cut -c1-300
$T ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_h.csv python scripts/run_forward_b64.py > gpurun_out/ncu.log 2>&1
