#!/bin/bash
T="timeout 250"
$T python -m pytest tests -m gpu -x -q 2>&1 | tail -8
$T python scripts/audit_configs.py --config 3 > gpurun_out/audit3.json 2> gpurun_out/audit3.err; tail -n 2 gpurun_out/audit3.json | cut -c1-330; tail -n 3 gpurun_out/audit3.err
$T python scripts/audit_configs.py --config 4 > gpurun_out/audit4.json 2> gpurun_out/audit4.err; cat gpurun_out/audit4.json | cut -c1-200
