#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-eager-gpu --no-config5"
show() { python -c "import sys,json; d=json.loads(open('$1').read()); r=d['roofline']; print(d['value'], d['ms_per_step'], 'enc_ms', r['kernel_ms_per_step'], 'frac', r['frac'], 'sus', d['sustained']['value'], 'e2e', d['e2e']['value'], d['e2e_track_u8']['value'])"; }
run() { tag=$1; shift; echo "$tag: $*"; env "$@" timeout 200 $B 2>/dev/null | tail -n 1 > gpurun_out/r2m_$tag.json; show gpurun_out/r2m_$tag.json; env "$@" LSD_TIMELINE=1 timeout 120 python scripts/run_forward_b64.py 2>&1 | grep timeline | tail -1; }
run base LSD_X=0
run a56h56 LSD_ART_CTAS=56 LSD_HF_CTAS=56
run a48h64 LSD_ART_CTAS=48 LSD_HF_CTAS=64
run a40h72 LSD_ART_CTAS=40 LSD_HF_CTAS=72
run a64h64 LSD_ART_CTAS=64 LSD_HF_CTAS=64
run a32h80 LSD_ART_CTAS=32 LSD_HF_CTAS=80
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
