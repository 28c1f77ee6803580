#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "u8 default"; timeout 120 python scripts/exp_graph_latency.py u8 2>&1 | tail -1
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-eager-gpu --no-config5"
timeout 200 $B 2>/dev/null | tail -n 1 > gpurun_out/r2r_bench.json; python -c "import sys,json; d=json.loads(open('gpurun_out/r2r_bench.json').read()); r=d['roofline']; print(d['value'], d['ms_per_step'], 'enc_ms', r['kernel_ms_per_step'], 'frac', r['frac'], 'sus', d['sustained']['value'], 'e2e', d['e2e'], d['e2e_fp32_upload']['value'], d['e2e_track_u8']['value'], d['latency'])"
