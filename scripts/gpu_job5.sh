#!/bin/bash
T="timeout 150"
$T python -m pytest tests -m gpu -x -q 2>&1 | tail -4
$T python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_now.json 2> gpurun_out/bench_now.err; cut -c1-250 gpurun_out/bench_now.json
LSD_UMMA_TRACE=2 $T python scripts/run_forward_b64.py 2> gpurun_out/trace13.log
LSD_UMMA_NTW=64 $T python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-120
$T ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_e.csv python scripts/run_forward_b64.py > gpurun_out/ncu.log 2>&1
