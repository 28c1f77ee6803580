#!/bin/bash
# fused stem + max-pool: bitwise check against the two-kernel path, GPU tests, bench before/after
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "scheduling_knobs" 2>&1 | tail -5
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
LSD_STEM_POOL_FUSE=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-eager-gpu --no-config5 2>/dev/null | tail -n 1 > gpurun_out/r2s_unfused.json; cut -c1-160 gpurun_out/r2s_unfused.json
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-eager-gpu --no-config5 2>/dev/null | tail -n 1 > gpurun_out/r2s_fused.json; cut -c1-160 gpurun_out/r2s_fused.json
python - <<'PY'
import json
for n in ("unfused","fused"):
    d=json.loads(open(f"gpurun_out/r2s_{n}.json").read())
    print(n, round(d["value"]), d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d["roofline"]["frac"], d["sustained"]["value"], d["e2e"]["value"])
PY
LSD_TIMELINE=1 timeout 120 python scripts/run_forward_b64.py 2>&1 | grep -i "M:\|S:\|T:" | head -40
