"""Preprocessed-mode evaluator (scripts/validate_pipeline.py:382-525 of the reference) on this backend: metrics against the
real reference's `compute_metrics` (tests/golden/validate_golden.json, made by tests/golden/make_validate_golden.py), and the
batching / resume / file logic with a scripted scorer (no GPU)."""
import json
import math
import os

import numpy as np
import pytest
import torch

import lipsync_b200 as lb
from lipsync_b200 import validate as val
from tests.conftest import ROOT


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "validate_golden.json")) as f:
        return json.load(f)


def _same(a, b):
    if isinstance(a, dict):
        return set(a) == set(b) and all(_same(a[k], b[k]) for k in a)
    if isinstance(a, float) and math.isnan(a):
        return isinstance(b, float) and math.isnan(b)
    return a == b


def test_metrics_match_reference(golden):
    assert len(golden) >= 6
    for name, g in golden.items():
        rows = [val.make_row(i, l, p) for i, (l, p) in enumerate(zip(g["labels"], g["probs"]))]
        assert _same(val.compute_metrics(rows), g["metrics"]), name


class _Dataset:
    """The reference dataset interface the evaluator needs: __len__, get_item(idx, train_mode_override=False)."""
    def __init__(self, n, missing=()):
        g = torch.Generator().manual_seed(3)
        self.v = torch.rand(n, 3, 4, 8, 8, generator=g)
        self.a = torch.rand(n, 1, 8, 16, generator=g)
        self.label = torch.randint(0, 2, (n,), generator=g)
        self.missing = set(missing)

    def __len__(self):
        return self.v.shape[0]

    def get_item(self, idx, train_mode_override=False):
        assert train_mode_override is False
        if idx in self.missing:
            return None
        return self.v[idx], self.a[idx], self.label[idx].float()


class _Scorer:
    """Stands in for Predictor.score_batches: logit = a fixed function of the sample."""
    def __init__(self):
        self.batch_sizes = []

    def score_batches(self, batches):
        out = []
        for v, a in batches:
            self.batch_sizes.append(v.shape[0])
            out.append((v.mean(dim=(1, 2, 3, 4)) - a.mean(dim=(1, 2, 3))) * 8.0)
        return out


def test_loop_rows_files_and_resume(tmp_path):
    ds = _Dataset(23, missing=(4, 17))
    sc = _Scorer()
    res = val.run_preprocessed_validation(ds, sc, output_dir=str(tmp_path), batch_size=5, save_every=10)
    rows = res["rows"]
    assert [r["sample_idx"] for r in rows] == [i for i in range(23) if i not in (4, 17)]
    assert sc.batch_sizes == [4, 5, 5, 4, 3]                       # skipped samples shrink their batch, like the reference
    for r in rows:
        i = r["sample_idx"]
        p = float(torch.sigmoid((ds.v[i].mean() - ds.a[i].mean()) * 8.0))
        assert abs(r["confidence"] - p) < 1e-6 and abs(r["manipulation_probability"] - (1 - p)) < 1e-6
        assert r["ground_truth"] == (0 if int(ds.label[i]) == 1 else 1)
        assert r["predicted_label"] == (0 if r["confidence"] >= 0.5 else 1)
        assert r["correct"] == int(r["predicted_label"] == r["ground_truth"])
    assert res["metrics"] == val.compute_metrics(rows)
    assert sorted(os.listdir(tmp_path)) == ["high_confidence_errors.csv", "metrics.json", "predictions.csv"]   # checkpoint removed
    assert val._read_csv(str(tmp_path / "predictions.csv")) == [dict(r) for r in rows]
    with open(tmp_path / "metrics.json") as f:
        assert _same(json.load(f), res["metrics"])
    # resume: everything is done already -> no scoring, same rows
    sc2 = _Scorer()
    res2 = val.run_preprocessed_validation(ds, sc2, output_dir=str(tmp_path), batch_size=5, resume=True)
    assert sc2.batch_sizes == [] and res2["rows"] == val._read_csv(str(tmp_path / "predictions.csv"))
    # resume from a partial predictions file: only the missing samples are scored
    val._write_csv(str(tmp_path / "predictions.csv"), rows[:9])
    sc3 = _Scorer()
    res3 = val.run_preprocessed_validation(ds, sc3, output_dir=str(tmp_path), batch_size=5, resume=True, n=20)
    assert sum(sc3.batch_sizes) == len([i for i in range(20) if i not in (4, 17)]) - 9
    assert sorted(r["sample_idx"] for r in res3["rows"]) == [i for i in range(20) if i not in (4, 17)]


def test_exports():
    assert lb.run_preprocessed_validation is val.run_preprocessed_validation and lb.compute_metrics is val.compute_metrics
