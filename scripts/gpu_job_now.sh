T="timeout 400"
$T python -m pytest tests -m gpu -q -x 2>&1 | tail -3
TAG="mean unrolled" $T python scripts/exp_knobs.py 2>&1 | tail -1
$T ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"planar_mean2|mha_core|layernorm|maxpool|video_rows" -c 60 --csv --log-file gpurun_out/launches_small.csv python scripts/run_forward_b64.py > /dev/null 2>&1
python - <<'PY'
import csv
lines=[l for l in open('gpurun_out/launches_small.csv') if not l.startswith('==')]
rows=[r for r in csv.DictReader(lines)]
for r in rows[-20:]:
    print(r['Kernel Name'][:40], r['Grid Size'], r['Metric Value'], r['Metric Unit'])
PY
