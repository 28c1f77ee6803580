T="timeout 400"
$T python -m pytest tests -m gpu -q -x 2>&1 | tail -3
TAG="token pieces parallel" $T python scripts/exp_knobs.py 2>&1 | tail -1
TAG="token serial" LSD_TOK_SERIAL=1 $T python scripts/exp_knobs.py 2>&1 | tail -1
LSD_TIMELINE=1 $T python scripts/run_forward_b64.py 2>&1 | grep timeline | tail -1
$T python scripts/audit_configs.py --config 4 2>/dev/null | tail -2 | cut -c1-120
