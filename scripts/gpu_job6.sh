#!/bin/bash
T="timeout 150"
$T python -m pytest tests -m gpu -x -q 2>&1 | tail -4
$T python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_now.json 2> gpurun_out/bench_now.err; cut -c1-250 gpurun_out/bench_now.json
LSD_UMMA_TRACE=2 $T python scripts/run_forward_b64.py 2> gpurun_out/trace14.log
$T ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_f.csv python scripts/run_forward_b64.py > gpurun_out/ncu.log 2>&1
$T python scripts/audit_configs.py --config 3 > gpurun_out/audit3.json 2> gpurun_out/audit3.err
$T python scripts/audit_configs.py --config 4 > gpurun_out/audit4.json 2> gpurun_out/audit4.err
timeout 300 python scripts/audit_configs.py --config 5 > gpurun_out/audit5.json 2> gpurun_out/audit5.err
tail -3 gpurun_out/audit3.json gpurun_out/audit4.json gpurun_out/audit5.json | cut -c1-400; tail -3 gpurun_out/audit*.err
