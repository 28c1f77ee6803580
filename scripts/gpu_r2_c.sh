#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -25
timeout 300 python scripts/dump_long_video_confs.py > gpurun_out/long_video_cuda_confs.json 2> gpurun_out/r2c_dump.err; echo "dump rc=$?"; tail -3 gpurun_out/r2c_dump.err
timeout 300 python - <<'PY'
import sys, json, torch
sys.path.insert(0, '.')
import bench
print(json.dumps(bench.eager_gpu_run(torch.device('cuda', 0), 64, 5, 3))[:700])
PY
