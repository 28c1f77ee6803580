#!/bin/bash
T="timeout 200"
$T ncu --set full --clock-control none --import-source on -k regex:planar_maxpool -c 1 -o gpurun_out/prof_maxpool2 -f python scripts/run_forward_b64.py > gpurun_out/ncu_full.log 2>&1
$T ncu --set full --clock-control none --import-source on -k regex:umma_conv_kernel -c 9 -o gpurun_out/prof_umma5 -f python scripts/run_forward_b64.py >> gpurun_out/ncu_full.log 2>&1
$T ncu --set full --clock-control none --import-source on -k regex:video_rows -c 1 -o gpurun_out/prof_vr2 -f python scripts/run_forward_b64.py >> gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
