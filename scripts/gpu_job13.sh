#!/bin/bash
T="timeout 150"
for n in 74 92 110 128 148; do
echo "side=$n"; LSD_SIDE_CTAS=$n $T python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e_track_u8']['value'])"
done
