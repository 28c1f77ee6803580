#!/bin/bash
run() { echo -n "$1: "; env $1 LSD_TIMELINE=1 timeout 120 python scripts/run_forward_b64.py 2>&1 | grep -i "timeline" | tail -1 | cut -c1-330; }
run X=1
run LSD_AUDIO_LATE=1
run LSD_AUDIO_AFTER_ROWS=1
run LSD_SIDE_CTAS=40
run LSD_SIDE_CTAS=24
run LSD_STEM_RING=0
