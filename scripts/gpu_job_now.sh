T="timeout 400"
$T python -m pytest tests -m gpu -q -x 2>&1 | tail -3
$T python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -n 1 > gpurun_out/bench_now.json
python -c "
import json; d=json.load(open('gpurun_out/bench_now.json')); print(d['value'], d['e2e']['value'], d['e2e_track_u8']['value'])"
$T python scripts/audit_configs.py --config 5 2>/dev/null | tail -1 | cut -c1-120
