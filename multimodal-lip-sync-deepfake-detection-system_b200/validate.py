"""Preprocessed-mode evaluator on the B200 backend: the model-only validation loop of the reference
(scripts/validate_pipeline.py:382-525 `_run_preprocessed_validation`, metrics :247-284 `compute_metrics`).

The reference scores `batch_size` precomputed (mouth-crop clip, log-mel) samples per `model(visual, audio)` call with a device
sync per batch; here the batches go through `Predictor.score_batches` (pinned host staging, H2D of batch k+1 under the forward
of batch k, lossless uint8 transport when the clips are uint8 / 255).  Rows, file names, label convention (manifest 1 = real,
0 = fake; metrics use 0 = real, 1 = fake with fake as the positive class) and the metrics dictionary are the reference's.
Dataset storage (Zarr / NPY / LMDB) stays the reference's `LipSyncDataset`: anything with `__len__` and
`get_item(idx, train_mode_override=False) -> (visual (3,T,H,W), audio (1,F,Ta), label) | None` can be passed in.
The PNG plots of the reference (matplotlib) are not produced.
"""
from __future__ import annotations

import csv
import json
import math
import os
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np
import torch

ROW_FIELDS = ["sample_idx", "ground_truth", "ground_truth_name", "predicted_label", "confidence", "manipulation_probability", "correct"]


def _avg_ranks(x: np.ndarray) -> np.ndarray:
    order = np.argsort(x, kind="mergesort")
    xs = x[order]
    ranks = np.empty(len(x), dtype=np.float64)
    i = 0
    while i < len(xs):
        j = i
        while j + 1 < len(xs) and xs[j + 1] == xs[i]:
            j += 1
        ranks[order[i:j + 1]] = 0.5 * (i + j) + 1.0
        i = j + 1
    return ranks


def compute_metrics(rows: Sequence[dict]) -> dict:
    """validate_pipeline.py:247-284 without sklearn / pandas: accuracy, precision / recall / F1 of the fake class
    (zero_division=0), FPR, FNR, ROC AUC of `1 - confidence` for the fake class (ties share their average rank), 2x2 confusion
    matrix.  With a single class present the AUC is NaN — what the reference computes with the scikit-learn of this image
    (tests/golden/validate_golden.json; older scikit-learn raised and the reference then reported 0.0)."""
    y_true = np.array([int(r["ground_truth"]) for r in rows], dtype=np.int64)
    y_pred = np.array([int(r["predicted_label"]) for r in rows], dtype=np.int64)
    y_score = np.array([float(r["confidence"]) for r in rows], dtype=np.float64)
    n = len(rows)
    tn = int(((y_true == 0) & (y_pred == 0)).sum()); fp = int(((y_true == 0) & (y_pred == 1)).sum())
    fn = int(((y_true == 1) & (y_pred == 0)).sum()); tp = int(((y_true == 1) & (y_pred == 1)).sum())
    accuracy = (tn + tp) / n if n else 0.0
    precision = tp / (tp + fp) if (tp + fp) > 0 else 0.0
    recall = tp / (tp + fn) if (tp + fn) > 0 else 0.0
    f1 = 2 * precision * recall / (precision + recall) if (precision + recall) > 0 else 0.0
    fpr = fp / (fp + tn) if (fp + tn) > 0 else 0.0
    fnr = fn / (fn + tp) if (fn + tp) > 0 else 0.0
    n_pos, n_neg = int((y_true == 1).sum()), int((y_true == 0).sum())
    if n_pos == 0 or n_neg == 0:
        roc_auc = float("nan")
    else:
        ranks = _avg_ranks(1.0 - y_score)
        roc_auc = (ranks[y_true == 1].sum() - n_pos * (n_pos + 1) / 2.0) / (n_pos * n_neg)
    return {
        "accuracy": round(float(accuracy), 6), "precision": round(float(precision), 6), "recall": round(float(recall), 6),
        "f1_score": round(float(f1), 6), "false_positive_rate": round(float(fpr), 6), "false_negative_rate": round(float(fnr), 6),
        "roc_auc": round(float(roc_auc), 6), "confusion_matrix": {"tn": tn, "fp": fp, "fn": fn, "tp": tp},
        "total_samples": n, "num_real": n_neg, "num_fake": n_pos,
    }


def make_row(sample_idx: int, gt_manifest: int, prob_real: float) -> dict:
    """validate_pipeline.py:470-487."""
    ground_truth = 0 if int(gt_manifest) == 1 else 1
    predicted_label = 0 if prob_real >= 0.5 else 1
    return {"sample_idx": int(sample_idx), "ground_truth": ground_truth, "ground_truth_name": "real" if ground_truth == 0 else "fake",
            "predicted_label": predicted_label, "confidence": float(prob_real), "manipulation_probability": 1.0 - float(prob_real),
            "correct": 1 if predicted_label == ground_truth else 0}


def _write_csv(path: str, rows: Sequence[dict]) -> None:
    with open(path, "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=ROW_FIELDS)
        w.writeheader()
        for r in rows:
            w.writerow({k: r[k] for k in ROW_FIELDS})


def _read_csv(path: str) -> List[dict]:
    out = []
    with open(path, newline="") as f:
        for r in csv.DictReader(f):
            out.append({"sample_idx": int(r["sample_idx"]), "ground_truth": int(r["ground_truth"]), "ground_truth_name": r["ground_truth_name"],
                        "predicted_label": int(r["predicted_label"]), "confidence": float(r["confidence"]),
                        "manipulation_probability": float(r["manipulation_probability"]), "correct": int(r["correct"])})
    return out


def run_preprocessed_validation(dataset, predictor, output_dir: Optional[str] = None, batch_size: int = 64, n: Optional[int] = None,
                                resume: bool = False, save_every: int = 0) -> Dict[str, object]:
    """Model-only validation over a preprocessed dataset (validate_pipeline.py:382-525).  `predictor`: a `lipsync_b200.Predictor`
    around the loaded model.  Returns {"rows": [...], "metrics": {...}} and, with `output_dir`, writes predictions.csv,
    high_confidence_errors.csv and metrics.json (plus predictions_checkpoint.csv every `save_every` samples, removed at the end)."""
    indices = list(range(len(dataset)))
    if n is not None:
        indices = indices[:n]
    rows: List[dict] = []
    ckpt = pred_path = None
    if output_dir is not None:
        os.makedirs(output_dir, exist_ok=True)
        ckpt, pred_path = os.path.join(output_dir, "predictions_checkpoint.csv"), os.path.join(output_dir, "predictions.csv")
        if resume:
            for path in (ckpt, pred_path):
                if os.path.isfile(path):
                    rows = _read_csv(path)
                    done = {r["sample_idx"] for r in rows}
                    indices = [i for i in indices if i not in done]
                    break

    def batches():
        """(indices, labels, visual batch, audio batch): samples the dataset cannot produce (None) are skipped, as in the reference."""
        for b0 in range(0, len(indices), batch_size):
            idx, vis, aud, lab = [], [], [], []
            for i in indices[b0:b0 + batch_size]:
                s = dataset.get_item(i, train_mode_override=False)
                if s is None:
                    continue
                v, a, l = s
                idx.append(i); vis.append(torch.as_tensor(v)); aud.append(torch.as_tensor(a)); lab.append(int(torch.as_tensor(l).reshape(-1)[0]))
            if idx:
                yield idx, lab, torch.stack(vis).contiguous(), torch.stack(aud).contiguous()

    meta: List[tuple] = []

    def feed():
        for idx, lab, v, a in batches():
            meta.append((idx, lab))
            yield v, a

    done_count = 0
    # score_batches returns when everything is scored; checkpoints are written per group of batches so that `save_every` still
    # bounds the work lost on interruption
    group = max(1, (save_every // batch_size) if save_every else 1 << 30)
    feeder = feed()
    while True:
        chunk = []
        for _ in range(group):
            nxt = next(feeder, None)
            if nxt is None:
                break
            chunk.append(nxt)
        if not chunk:
            break
        m0 = len(meta) - len(chunk)
        logits = predictor.score_batches(chunk)
        for (idx, lab), lg in zip(meta[m0:], logits):
            probs = torch.sigmoid(lg.float()).numpy()
            for j, i in enumerate(idx):
                rows.append(make_row(i, lab[j], float(probs[j])))
                done_count += 1
        if ckpt is not None and save_every and done_count > 0:
            _write_csv(ckpt, rows)
    metrics = compute_metrics(rows) if rows else {}
    if output_dir is not None:
        _write_csv(pred_path, rows)
        if os.path.isfile(ckpt):
            try:
                os.unlink(ckpt)
            except OSError:
                pass
        _write_csv(os.path.join(output_dir, "high_confidence_errors.csv"),
                   [r for r in rows if r["correct"] == 0 and (r["confidence"] > 0.9 or r["manipulation_probability"] > 0.9)])
        with open(os.path.join(output_dir, "metrics.json"), "w") as f:
            json.dump(metrics, f, indent=2)
    return {"rows": rows, "metrics": metrics}
