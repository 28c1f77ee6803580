#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/check_tok_fused.py > gpurun_out/r2d_tokfused.log 2>&1; echo "tokfused rc=$?"; tail -16 gpurun_out/r2d_tokfused.log
timeout 120 python scripts/trace_tok_fused.py 64 > gpurun_out/r2d_tokf_trace.log 2>&1; echo "trace rc=$?"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-eager-gpu --no-config5"
for c in 0 7 15; do echo "STEM_CHUNK=$c"; LSD_STEM_CHUNK=$c timeout 300 $B 2>/dev/null | tail -n 1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print(d['value'], d['ms_per_step'], 'enc_ms', r['kernel_ms_per_step'], 'frac', r['frac'], 'sus', d['sustained']['value'])"; done
for a in 90 110; do echo "ART_CTAS=$a"; LSD_ART_CTAS=$a timeout 300 $B 2>/dev/null | tail -n 1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print(d['value'], d['ms_per_step'], 'enc_ms', r['kernel_ms_per_step'], 'frac', r['frac'])"; done
LSD_TIMELINE=1 timeout 120 python scripts/run_forward_b64.py 2>&1 | grep timeline | tail -1
