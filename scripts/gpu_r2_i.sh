#!/bin/bash
mkdir -p gpurun_out
timeout 180 python -u scripts/check_tok_front.py > gpurun_out/r2i_tokfront.log 2>&1; echo "tokfront rc=$?"; tail -22 gpurun_out/r2i_tokfront.log
