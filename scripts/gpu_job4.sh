#!/bin/bash
T="timeout 150"
$T python -m pytest tests -m gpu -x -q 2>&1 | tail -4
$T python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_now.json 2> gpurun_out/bench_now.err; cut -c1-250 gpurun_out/bench_now.json
LSD_UMMA_TRACE=2 $T python scripts/run_forward_b64.py 2> gpurun_out/trace12.log
LSD_UMMA_NT3=256 LSD_UMMA_NT4=256 LSD_UMMA_TRACE=1 $T python scripts/run_forward_b64.py 2> gpurun_out/trace12_nt256.log
$T ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_d.csv python scripts/run_forward_b64.py > gpurun_out/ncu.log 2>&1
$T ncu --set full --clock-control none --import-source on -k regex:planar_maxpool -c 1 -o gpurun_out/prof_maxpool -f python scripts/run_forward_b64.py > gpurun_out/ncu_full.log 2>&1
