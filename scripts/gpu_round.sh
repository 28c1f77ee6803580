#!/bin/bash
# One GPU round: parity tests, bench line, ncu launch lists (times; DRAM bytes + tensor pipe), ncu --set full of the visual
# encoder launches, sub-path audits.  Everything lands in gpurun_out/ (copy what should be judged into profiles/).
T="timeout 400"
$T python -m pytest tests -m gpu -x -q 2>&1 | tail -3
$T python bench.py --steps 20 --warmup 3 2> gpurun_out/bench_final.err | tail -n 1 > gpurun_out/bench_final.json; cut -c1-200 gpurun_out/bench_final.json
$T ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
$T ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -c 400 --csv --log-file gpurun_out/launches_traffic.csv python scripts/run_forward_b64.py > gpurun_out/ncu_traffic.log 2>&1
LSD_AUDIO_LATE=1 $T ncu --set full --clock-control none --import-source on -k regex:umma_conv_kernel -c 9 -o gpurun_out/prof_umma7 -f python scripts/run_forward_b64.py > gpurun_out/ncu_full.log 2>&1
$T python scripts/audit_configs.py --config 3 > gpurun_out/audit3.json 2>/dev/null
$T python scripts/audit_configs.py --config 4 > gpurun_out/audit4.json 2>/dev/null
$T python scripts/audit_configs.py --config 5 > gpurun_out/audit5_n1.json 2>/dev/null; cut -c1-200 gpurun_out/audit5_n1.json
