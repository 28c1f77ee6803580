#!/bin/bash
for sk in 0 16 32 48; do
echo "skip=$sk"; LSD_UMMA_SKIP=$sk LSD_UMMA_TRACE=1 timeout 120 python scripts/run_forward_b64.py 2>&1 | grep "visual_encoder.stem" | tail -1
done
