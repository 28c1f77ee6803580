#!/bin/bash
T="timeout 250"
$T python -m pytest tests -m gpu -x -q 2>&1 | tail -6
$T python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_now.json 2> gpurun_out/bench_now.err; python -c "
import json; d=json.load(open('gpurun_out/bench_now.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e_track_u8']['value'], d['roofline']['frac'])"
$T python scripts/audit_configs.py --config 5 | cut -c1-300
LSD_NO_PIPELINE=1 $T python scripts/audit_configs.py --config 5 | cut -c1-300
