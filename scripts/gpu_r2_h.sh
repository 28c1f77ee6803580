#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-eager-gpu --no-config5"
show() { python -c "import sys,json; d=json.loads(open('$1').read()); r=d['roofline']; print(d['value'], d['ms_per_step'], 'enc_ms', r['kernel_ms_per_step'], 'frac', r['frac'], 'sus', d['sustained']['value'], 'e2e', d['e2e']['value'], d['e2e_track_u8']['value'], d['latency']['infer_confidence_ms']['best'], d['latency'].get('graph_captures'))"; }
run() { tag=$1; shift; echo "$tag: $*"; env "$@" timeout 200 $B 2>/dev/null | tail -n 1 > gpurun_out/r2h_$tag.json; show gpurun_out/r2h_$tag.json; }
run base LSD_X=0
run side16 LSD_SIDE_CTAS=16 LSD_ART_CTAS=74
run side32 LSD_SIDE_CTAS=32 LSD_ART_CTAS=74
run side8 LSD_SIDE_CTAS=8 LSD_ART_CTAS=74
run audiolate LSD_AUDIO_LATE=1
run side16dyn LSD_SIDE_CTAS=16 LSD_ART_CTAS=74 LSD_UMMA_DYNAMIC=1
run side16rows LSD_SIDE_CTAS=16 LSD_ART_CTAS=74 LSD_AUDIO_AFTER_ROWS=1
LSD_SIDE_CTAS=16 LSD_ART_CTAS=74 LSD_TIMELINE=1 timeout 120 python scripts/run_forward_b64.py 2>&1 | grep timeline | tail -1
