timeout 600 python -m pytest tests -m gpu -q -x -k "scheduling_knobs" 2>&1 | tail -5
