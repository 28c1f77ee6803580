T="timeout 300"
for m in 116 124 132 140; do TAG="main=$m" LSD_MAIN_CTAS=$m $T python scripts/exp_knobs.py 2>&1 | tail -1; done
for m in 124 132; do TAG="audio-first main=$m" LSD_AUDIO_FIRST=1 LSD_MAIN_CTAS=$m $T python scripts/exp_knobs.py 2>&1 | tail -1; done
LSD_MAIN_CTAS=132 LSD_TIMELINE=1 $T python scripts/run_forward_b64.py 2>&1 | grep timeline | tail -1
