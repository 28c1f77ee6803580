#!/bin/bash
T="timeout 250"
for v in default 2; do
if [ $v != default ]; then export LSD_UMMA_ISSUERS=$v; fi
echo "issuers=$v"
LSD_UMMA_TRACE=1 $T python scripts/run_forward_b64.py 2>&1 | grep "^\[umma\] visual_encoder\|^\[umma\] art" | tail -15 | awk '{print $2, $3, $4, "total", $(NF-1)}'
done
unset LSD_UMMA_ISSUERS
$T python -m pytest tests -m gpu -x -q 2>&1 | tail -3
$T python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -n 1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e_track_u8']['value'], d['roofline']['frac'])"
