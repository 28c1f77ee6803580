#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-eager-gpu --no-config5"
show() { python -c "import sys,json; d=json.loads(open('$1').read()); r=d['roofline']; print(d['value'], d['ms_per_step'], 'enc_ms', r['kernel_ms_per_step'], 'frac', r['frac'], 'sus', d['sustained']['value'], 'e2e', d['e2e']['value'], d['e2e_track_u8']['value'], d['gpu_launches'])"; }
run() { tag=$1; shift; echo "$tag: $*"; env "$@" timeout 200 $B 2>/dev/null | tail -n 1 > gpurun_out/r2n_$tag.json; show gpurun_out/r2n_$tag.json; env "$@" LSD_TIMELINE=1 timeout 120 python scripts/run_forward_b64.py 2>&1 | grep timeline | tail -1; }
run base LSD_X=0
run a40h72 LSD_ART_CTAS=40 LSD_HF_CTAS=72
run a56h90 LSD_ART_CTAS=56 LSD_HF_CTAS=90
run hfser LSD_HF_SERIAL=1
