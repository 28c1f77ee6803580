timeout 300 python scripts/exp_clocks.py 2>&1 | tail -16
