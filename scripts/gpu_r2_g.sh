#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-eager-gpu --no-config5"
show() { python -c "import sys,json; d=json.loads(open('$1').read()); r=d['roofline']; print(d['value'], d['ms_per_step'], 'enc_ms', r['kernel_ms_per_step'], 'frac', r['frac'], 'sus', d['sustained']['value'], 'e2e', d['e2e']['value'], d['e2e_track_u8']['value'])"; }
echo "HF parallel (default)"; timeout 200 $B 2>/dev/null | tail -n 1 > gpurun_out/r2g_bench_hfpar.json; show gpurun_out/r2g_bench_hfpar.json
echo "HF serial"; LSD_HF_SERIAL=1 timeout 200 $B 2>/dev/null | tail -n 1 > gpurun_out/r2g_bench_hfser.json; show gpurun_out/r2g_bench_hfser.json
for c in 50 100; do echo "HF_CTAS=$c"; LSD_HF_CTAS=$c timeout 200 $B 2>/dev/null | tail -n 1 > gpurun_out/r2g_bench_hf$c.json; show gpurun_out/r2g_bench_hf$c.json; done
echo "HF early"; LSD_HF_EARLY=1 timeout 200 $B 2>/dev/null | tail -n 1 > gpurun_out/r2g_bench_hfearly.json; show gpurun_out/r2g_bench_hfearly.json
LSD_TIMELINE=1 timeout 120 python scripts/run_forward_b64.py 2>&1 | grep timeline | tail -1
LSD_HF_SERIAL=1 LSD_TIMELINE=1 timeout 120 python scripts/run_forward_b64.py 2>&1 | grep timeline | tail -1
timeout 120 python scripts/exp_latency.py 2>&1 | tail -3
LSD_TOK_FUSED=0 timeout 120 python scripts/exp_latency.py 2>&1 | tail -3
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
