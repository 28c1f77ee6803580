#!/bin/bash
# tuning experiments: per-layer cycle totals of the tcgen05 kernel under different tile configurations
T="timeout 150"
$T python -m pytest tests -m gpu -x -q 2>&1 | tail -4
$T python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_now.json 2> gpurun_out/bench_now.err; cut -c1-250 gpurun_out/bench_now.json
LSD_UMMA_TRACE=2 $T python scripts/run_forward_b64.py 2> gpurun_out/trace11.log
LSD_UMMA_NT3=128 LSD_UMMA_NT4=128 LSD_UMMA_TRACE=1 $T python scripts/run_forward_b64.py 2> gpurun_out/trace11_nt128.log
LSD_UMMA_WKB=20 LSD_UMMA_TRACE=1 $T python scripts/run_forward_b64.py 2> gpurun_out/trace11_wkb20.log
LSD_UMMA_NT3=128 LSD_UMMA_NT4=128 LSD_UMMA_WKB=20 LSD_UMMA_TRACE=1 $T python scripts/run_forward_b64.py 2> gpurun_out/trace11_nt128_wkb20.log
LSD_UMMA_NT3=128 LSD_UMMA_NT4=128 $T python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-120
LSD_UMMA_WKB=20 $T python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-120
$T python scripts/audit_configs.py --config pcie
$T ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c.csv python scripts/run_forward_b64.py > gpurun_out/ncu.log 2>&1
