#!/bin/bash
mkdir -p gpurun_out
for c in 0 1; do
echo "cta2=$c"; LSD_UMMA_CTA2=$c LSD_UMMA_TRACE=1 timeout 120 python scripts/run_forward_b64.py 2>&1 | grep "visual_encoder.stem\|visual_encoder.layer1" | tail -3 | cut -c1-60,130-300
done
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
LSD_UMMA_CTA2=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-eager-gpu --no-config5 2>/dev/null | tail -n 1 > gpurun_out/r2v_single.json
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-eager-gpu --no-config5 2>/dev/null | tail -n 1 > gpurun_out/r2v_pairs.json
python - <<'PY'
import json
for n in ("single","pairs"):
    d=json.loads(open(f"gpurun_out/r2v_{n}.json").read())
    print(n, round(d["value"]), d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d["roofline"]["frac"], d["sustained"]["value"], d["e2e"]["value"], d["e2e_track_u8"]["value"])
PY
LSD_TIMELINE=1 timeout 120 python scripts/run_forward_b64.py 2>&1 | grep -i "timeline" | tail -1
