#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "u8_transport or preprocessed" 2>&1 | tail -4
for i in 1 2; do timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-eager-gpu --no-config5 2>/dev/null | tail -n 1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['e2e']['ms_per_step'], d['e2e']['transport'], d['e2e']['h2d_bytes_per_step'], round(d['e2e_fp32_upload']['value']), round(d['e2e_track_u8']['value']))"; done
