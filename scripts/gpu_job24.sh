#!/bin/bash
T="timeout 120"
for sk in 0 4 8 12; do
echo "skip=$sk"
LSD_UMMA_SKIP=$sk LSD_UMMA_TRACE=1 $T python scripts/run_forward_b64.py 2>&1 | grep "visual_encoder.stem \|art.hf0\|layer1.conv1" | tail -3 | cut -c1-50,150-240
done
