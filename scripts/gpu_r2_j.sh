#!/bin/bash
mkdir -p gpurun_out
timeout 60 python -u scripts/trace_tok_fused.py 64 > gpurun_out/r2j_tok_trace.log 2>&1; echo "trace rc=$?"
grep tokfront gpurun_out/r2j_tok_trace.log | cut -c1-3000
