#!/bin/bash
T="timeout 250"
LSD_UMMA_TRACE=2 $T python scripts/run_forward_b64.py 2> gpurun_out/trace15.log; grep -A1 "t0.in  \|t0.ff2" gpurun_out/trace15.log | head -4 | cut -c1-260
$T python -m pytest tests -m gpu -x -q 2>&1 | tail -3
$T python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -n 1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e_track_u8']['value'], d['roofline']['frac'])"
