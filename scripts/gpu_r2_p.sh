#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "stem_ring" 2>&1 | tail -4
LSD_TIMELINE=1 timeout 120 python scripts/run_forward_b64.py 2>&1 | grep -i "timeline" | tail -1 | cut -c1-200
LSD_STEM_POOL_INLINE=0 LSD_TIMELINE=1 timeout 120 python scripts/run_forward_b64.py 2>&1 | grep -i "timeline" | tail -1 | cut -c1-200
