#!/bin/bash
mkdir -p gpurun_out
timeout 120 python scripts/trace_tok_fused.py 64 > gpurun_out/r2b_tokf_trace.log 2>&1; echo "trace rc=$?"
timeout 600 python bench.py --steps 20 --warmup 3 2> gpurun_out/r2b_bench.err | tail -n 1 > gpurun_out/r2b_bench.json; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2b_bench.json; tail -5 gpurun_out/r2b_bench.err
for a in 90 100 110; do LSD_ART_CTAS=$a timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-eager-gpu --no-config5 2> /dev/null | tail -n 1 | cut -c1-140; done
