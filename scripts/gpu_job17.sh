#!/bin/bash
T="timeout 250"
$T python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for v in 0 1; do
if [ $v = 1 ]; then export LSD_AUDIO_LATE=1; fi
$T python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -n 1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('late' if '$v'=='1' else 'early', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e_track_u8']['value'], d['roofline']['frac'])"
done
unset LSD_AUDIO_LATE
$T python scripts/audit_configs.py --config 5 | cut -c1-200
