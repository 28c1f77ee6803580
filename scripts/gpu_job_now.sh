T="timeout 300"
$T python -m pytest tests -m gpu -q -x 2>&1 | tail -3
TAG="mean rewrite" $T python scripts/exp_knobs.py 2>&1 | tail -1
for a in 86 98 110; do TAG="side=74 art=$a" LSD_ART_CTAS=$a $T python scripts/exp_knobs.py 2>&1 | tail -1; done
for a in 74 98; do TAG="side=60 art=$a" LSD_SIDE_CTAS=60 LSD_ART_CTAS=$a $T python scripts/exp_knobs.py 2>&1 | tail -1; done
LSD_TIMELINE=1 $T python scripts/run_forward_b64.py 2>&1 | grep timeline | tail -1
