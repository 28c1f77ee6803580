#!/bin/bash
# Round-2 profile pass: bench line, ncu launch list of the same command, ncu metrics pass (time, DRAM bytes, tensor pipe) of one
# B=64 forward, ncu --set full of the visual-encoder launches and of the two fused token kernels, sub-path audits.
mkdir -p gpurun_out
T="timeout 400"
TAG=${1:-r02}
$T python bench.py --steps 20 --warmup 3 2> gpurun_out/${TAG}_bench.err | tail -n 1 > gpurun_out/${TAG}_bench_n1.json; cut -c1-300 gpurun_out/${TAG}_bench_n1.json
$T ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/${TAG}_launches_bench_py.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-eager-gpu --no-config5 > gpurun_out/${TAG}_ncu_bench.log 2>&1
$T ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_traffic_tensorpipe.csv python scripts/run_forward_b64.py > gpurun_out/${TAG}_ncu_traffic.log 2>&1
LSD_AUDIO_LATE=1 $T ncu --set full --clock-control none --import-source on -k "regex:umma_conv_kernel|stem_ring_kernel" -c 9 -o gpurun_out/${TAG}_prof_umma_encoder -f python scripts/run_forward_b64.py > gpurun_out/${TAG}_ncu_full.log 2>&1
$T ncu --set full --clock-control none --import-source on -k regex:tok_f -c 2 -o gpurun_out/${TAG}_prof_tok -f python scripts/run_forward_b64.py > gpurun_out/${TAG}_ncu_full_tok.log 2>&1
# glue kernels of the critical path + the log-mel FFT kernel (pipe utilisation: what bounds them)
$T ncu --set full --clock-control none --import-source on -k "regex:video_rows_tma_kernel|planar_maxpool_direct_kernel" -c 2 -o gpurun_out/${TAG}_prof_glue -f python scripts/run_forward_b64.py > gpurun_out/${TAG}_ncu_full_glue.log 2>&1
$T ncu --set full --clock-control none --import-source on -k regex:logmel_fft_kernel -s 4 -c 1 -o gpurun_out/${TAG}_prof_logmel -f python scripts/audit_configs.py --config 3 --batches 512 > gpurun_out/${TAG}_ncu_full_logmel.log 2>&1
$T python scripts/audit_configs.py --config 3 > gpurun_out/${TAG}_audit3.json 2>/dev/null
$T python scripts/audit_configs.py --config 4 > gpurun_out/${TAG}_audit4.json 2>/dev/null; cut -c1-400 gpurun_out/${TAG}_audit4.json
$T python scripts/audit_configs.py --config 5 > gpurun_out/${TAG}_audit5_n1.json 2>/dev/null; cut -c1-300 gpurun_out/${TAG}_audit5_n1.json
ls -la gpurun_out/${TAG}_*
