#!/bin/bash
# round 2, run A: first look at the fused transformer kernel + baseline experiments
mkdir -p gpurun_out
timeout 300 python scripts/check_tok_fused.py > gpurun_out/r2a_tokfused.log 2>&1; echo "tokfused rc=$?"; tail -25 gpurun_out/r2a_tokfused.log
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2> gpurun_out/r2a_bench.err | tail -n 1 > gpurun_out/r2a_bench.json; cut -c1-400 gpurun_out/r2a_bench.json
LSD_TOK_FUSED=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2> /dev/null | tail -n 1 > gpurun_out/r2a_bench_unfused.json; cut -c1-200 gpurun_out/r2a_bench_unfused.json
LSD_AUDIO_LATE=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2> /dev/null | tail -n 1 > gpurun_out/r2a_bench_audiolate.json; cut -c1-200 gpurun_out/r2a_bench_audiolate.json
LSD_SIDE_CTAS=32 LSD_ART_CTAS=100 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2> /dev/null | tail -n 1 > gpurun_out/r2a_bench_side32.json; cut -c1-200 gpurun_out/r2a_bench_side32.json
LSD_SIDE_CTAS=74 LSD_ART_CTAS=120 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2> /dev/null | tail -n 1 > gpurun_out/r2a_bench_art120.json; cut -c1-200 gpurun_out/r2a_bench_art120.json
LSD_TIMELINE=1 timeout 120 python scripts/run_forward_b64.py 2>&1 | grep timeline | tail -2
