#!/bin/bash
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
LSD_TOKF_TRACE=1 timeout 120 python scripts/run_forward_b64.py 2>&1 | grep "\[tokf\]" | tail -2 | cut -c1-6000
