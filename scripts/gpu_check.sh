#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
LSD_TIMELINE=1 timeout 120 python scripts/run_forward_b64.py 2>&1 | grep -i "timeline" | tail -1
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-eager-gpu --no-config5 2>/dev/null | tail -n 1 > gpurun_out/check_bench.json
python - <<'PY'
import json
d=json.loads(open("gpurun_out/check_bench.json").read())
print(round(d["value"]), d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d["roofline"]["frac"], d["sustained"]["value"], d["e2e"]["value"], d["e2e"]["transport"], d["e2e_track_u8"]["value"])
PY
