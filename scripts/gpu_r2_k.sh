#!/bin/bash
mkdir -p gpurun_out
timeout 180 python -u scripts/check_tok_front.py > gpurun_out/r2k_tokfront.log 2>&1; echo "tokfront rc=$?"; tail -20 gpurun_out/r2k_tokfront.log
timeout 120 python -u scripts/check_tok_fused.py > gpurun_out/r2k_tokfused.log 2>&1; echo "tokfused rc=$?"; tail -6 gpurun_out/r2k_tokfused.log
timeout 60 python -u scripts/trace_tok_fused.py 64 > gpurun_out/r2k_tok_trace.log 2>&1; echo "trace rc=$?"
grep tokfront gpurun_out/r2k_tok_trace.log | cut -c1-1500
