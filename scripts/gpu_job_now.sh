T="timeout 400"
$T python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for kb in 64 96; do TAG="stage_kb=$kb" LSD_UMMA_STAGE_KB=$kb $T python scripts/exp_knobs.py 2>&1 | tail -1; done
