"""Known answers of the reference's preprocessed-mode evaluator (scripts/validate_pipeline.py:247-284 `compute_metrics`, and the
row construction of `_run_preprocessed_validation`, :470-487), produced by the REAL reference functions in this container.
librosa / matplotlib are not installed here and are stubbed (neither is touched by these functions).

    python tests/golden/make_validate_golden.py   ->  tests/golden/validate_golden.json
"""
import importlib.util
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference():
    m = types.ModuleType("matplotlib"); m.use = lambda *a, **k: None
    pp = types.ModuleType("matplotlib.pyplot"); m.pyplot = pp
    sys.modules.setdefault("matplotlib", m); sys.modules.setdefault("matplotlib.pyplot", pp)
    sys.modules.setdefault("librosa", types.ModuleType("librosa"))
    sys.path.insert(0, "/root/reference"); sys.path.insert(0, "/root/reference/scripts")
    spec = importlib.util.spec_from_file_location("ref_validate_pipeline", "/root/reference/scripts/validate_pipeline.py")
    vp = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(vp)
    return vp


def scenarios():
    """(name, manifest labels (1 = real, 0 = fake), P(real) per sample)"""
    rng = np.random.default_rng(11)
    out = []
    n = 200
    lab = rng.integers(0, 2, n)
    p = np.clip(np.where(lab == 1, 0.7, 0.3) + 0.25 * rng.standard_normal(n), 0.0, 1.0)
    out.append(("mixed_200", lab.tolist(), p.tolist()))
    out.append(("all_real", [1] * 7, rng.uniform(0.2, 1.0, 7).tolist()))
    out.append(("all_fake_all_wrong", [0] * 5, rng.uniform(0.6, 1.0, 5).tolist()))
    out.append(("ties_at_half", [1, 0, 1, 0, 1, 0], [0.5, 0.5, 0.5, 0.25, 0.75, 0.5]))
    lab = rng.integers(0, 2, 33)
    out.append(("coarse_scores", lab.tolist(), (rng.integers(0, 5, 33) / 4.0).tolist()))
    out.append(("single", [0], [0.1]))
    return out


def main():
    import pandas as pd
    vp = load_reference()
    gold = {}
    for name, labels, probs in scenarios():
        rows = []
        for i, (gt_manifest, prob_real) in enumerate(zip(labels, probs)):      # validate_pipeline.py:470-487
            ground_truth = 0 if int(gt_manifest) == 1 else 1
            predicted_label = 0 if float(prob_real) >= 0.5 else 1
            rows.append({"sample_idx": i, "ground_truth": ground_truth, "ground_truth_name": "real" if ground_truth == 0 else "fake",
                         "predicted_label": predicted_label, "confidence": float(prob_real), "manipulation_probability": 1.0 - float(prob_real),
                         "correct": 1 if predicted_label == ground_truth else 0})
        gold[name] = {"labels": labels, "probs": probs, "metrics": vp.compute_metrics(pd.DataFrame(rows))}
    with open(os.path.join(HERE, "validate_golden.json"), "w") as f:
        json.dump(gold, f, indent=1)
    print("wrote", len(gold), "scenarios")


if __name__ == "__main__":
    main()
