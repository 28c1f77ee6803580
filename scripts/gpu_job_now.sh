timeout 200 python -m pytest tests -m gpu -q -x -k "logits or composition or odd_shapes or subpath" 2>&1 | tail -3
