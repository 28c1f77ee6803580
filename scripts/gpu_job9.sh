#!/bin/bash
T="timeout 200"
$T python -m pytest tests -m gpu -x -q 2>&1 | tail -6
$T python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_now.json 2> gpurun_out/bench_now.err; cut -c1-250 gpurun_out/bench_now.json
$T python scripts/audit_configs.py --config 3 > gpurun_out/audit3.json 2> gpurun_out/audit3.err; tail -n 3 gpurun_out/audit3.json | cut -c1-330; tail -n 3 gpurun_out/audit3.err
$T ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_g.csv python scripts/run_forward_b64.py > gpurun_out/ncu.log 2>&1
