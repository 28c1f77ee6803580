#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -u scripts/check_tok_fused.py > gpurun_out/r2l_tokfused.log 2>&1; echo "tokfused rc=$?"; tail -22 gpurun_out/r2l_tokfused.log
timeout 60 python -u scripts/trace_tok_fused.py 64 > gpurun_out/r2l_tok_trace.log 2>&1; echo "trace rc=$?"
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-eager-gpu --no-config5"
timeout 200 $B 2>/dev/null | tail -n 1 > gpurun_out/r2l_bench.json; python -c "import sys,json; d=json.loads(open('gpurun_out/r2l_bench.json').read()); r=d['roofline']; print(d['value'], d['ms_per_step'], 'enc_ms', r['kernel_ms_per_step'], 'frac', r['frac'], 'sus', d['sustained']['value'], 'e2e', d['e2e']['value'], d['e2e_track_u8']['value'], d['gpu_launches'], d['latency'])"
LSD_TIMELINE=1 timeout 120 python scripts/run_forward_b64.py 2>&1 | grep timeline | tail -1
