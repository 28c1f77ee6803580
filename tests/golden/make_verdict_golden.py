"""Golden decisions from the REAL reference `_predict_long_video` (imported from /root/reference, this container only).

    python tests/golden/make_verdict_golden.py

The aggregation / gate block (app/inference/predictor.py:856-1155, verdict :1235) is inline code of
`Predictor._predict_long_video`; it is exercised end to end here on synthetic single-track videos with the three
preprocessing names monkeypatched and `_infer_confidence` scripted to return a prescribed confidence
per window.  For every scenario the fixture stores what the reference fed to the block (per-window confidence,
speaking activity, VAD coverage, mouth-motion result, thresholds) and everything it decided.
Output (committed): tests/golden/verdict_golden.json
"""
import asyncio
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
sys.modules.setdefault("librosa", types.ModuleType("librosa"))

import torch  # noqa: E402
import app.inference.predictor as P  # noqa: E402

KEYS = ["verdict", "is_real", "is_fake", "confidence", "window_weighted_confidence", "window_fake_vote_ratio",
        "window_consensus_uncertain", "strict_fake_evidence", "sparse_real_guard_applied", "mouth_motion_override_applied",
        "override_reason", "temporal_confidence_drop", "temporal_drift", "first_half_avg_confidence", "second_half_avg_confidence"]


def scenarios():
    rng = np.random.RandomState(7)
    out = {}
    n = 40
    out["all_real"] = np.clip(0.8 + 0.1 * rng.randn(n), 0, 1)
    out["all_fake"] = np.clip(0.15 + 0.05 * rng.randn(n), 0, 1)
    c = np.clip(0.85 + 0.05 * rng.randn(n), 0, 1); c[25:35] = 0.05 + 0.02 * rng.rand(10); out["localised_fake_segment"] = c
    c = np.clip(0.8 + 0.05 * rng.randn(n), 0, 1); c[[3, 17]] = 0.1; out["two_noise_windows"] = c
    c = np.clip(0.3 + 0.05 * rng.randn(n), 0, 1); c[:3] = 0.9; c[20:23] = 0.92; out["mostly_fake_some_strong_real"] = c
    c = np.full(n, 0.02); c[10] = 0.45; out["sparse_real_signal"] = c
    c = np.clip(0.9 - 0.02 * np.arange(n) + 0.01 * rng.randn(n), 0, 1); out["temporal_drift"] = c
    out["borderline"] = np.clip(0.5 + 0.03 * rng.randn(n), 0, 1)
    out["short_three_windows"] = np.asarray([0.7, 0.2, 0.6])
    c = np.clip(0.25 + 0.05 * rng.randn(n), 0, 1); c[5:8] = 0.8; c[30:32] = 0.7; out["mixed_consensus"] = c
    c = np.full(24, 0.45); c[::2] = 0.2; out["long_fake_run"] = np.concatenate([np.full(10, 0.1), c])
    return {k: [float(x) for x in v] for k, v in out.items()}


def run(confs, motion, seed, mouth=None, **pred_kwargs):
    n = len(confs)
    rng = np.random.RandomState(seed)
    n_frames = 32 + 8 * (n - 1)
    # synthetic single track: smooth random crops; `motion` scales the frame-to-frame change (mouth-motion check)
    base = rng.rand(3, 1, 96, 96).astype(np.float32)
    drift = np.cumsum(motion * rng.randn(3, n_frames, 96, 96).astype(np.float32), axis=1)
    crops = np.clip(base + drift, 0, 1).astype(np.float32)
    starts = [8 * i for i in range(n)]
    chunks = [np.ascontiguousarray(crops[:, s:s + 32]) for s in starts]
    ta = int(n_frames / 15.0 * 100.0)
    mel = (-80.0 * rng.rand(1, 80, ta)).astype(np.float32)
    vad = rng.rand(ta) > 0.3
    track = {"track_id": 0, "chunks": chunks, "chunk_starts": starts, "stability": 0.9, "hits": n_frames,
             "consecutive_miss_max": 0, "track_start_frame": 0, "track_end_frame": n_frames - 1}
    P.preprocess_video_tracks_chunked = lambda *a, **k: ([track], 15.0, n_frames)
    P.preprocess_audio = lambda *a, **k: mel
    P.detect_voice_activity = lambda *a, **k: (vad, ta / 100.0)
    p = P.Predictor.__new__(P.Predictor)
    defaults = dict(confidence_threshold=0.5, uncertainty_margin=0.05, confidence_smoothing="median", trim_ratio=0.1,
                    max_tracks=6, chunk_size=32, chunk_stride=8, max_total_frames=None, confidence_margin=0.10,
                    mouth_motion_check_enabled=True, mouth_motion_low_threshold=0.015, mouth_motion_fake_penalty=0.10,
                    audio_energy_high_threshold=-25.0, audio_energy_low_threshold=-50.0, weak_real_gate=0.08,
                    weak_real_window_threshold=0.30, fake_vote_gate=0.15, fake_vote_min_windows=5,
                    refine_margin=0.08, refine_top_k=2, long_video_threshold_sec=3.0, device=torch.device("cpu"))
    defaults.update(pred_kwargs)
    for k, v in defaults.items():
        setattr(p, k, v)
    it = iter(confs)
    p._infer_confidence = lambda v, a: next(it)
    if mouth is not None:   # script the mouth-motion helper's outcome (the gate block itself stays the reference's)
        p._aggregate_mouth_motion_check = lambda *a, **k: {"check_result": mouth, "audio_energy": -20.0, "mouth_motion_energy": 0.001,
                                                           "samples_checked": 5, "counts": {mouth: 5}}
    res = asyncio.run(p._predict_long_video(None, None, 0.0))
    rec = {k: res[k] for k in KEYS}
    rec["inputs"] = {
        "window_confs": [float(w["confidence"]) for w in res["window_results"]],
        "window_speaking": [float(w["speaking_activity"]) for w in res["window_results"]],
        "window_vad": [float(w["vad_coverage"]) for w in res["window_results"]],
        "mouth_check_result": res["mouth_motion_check"]["check_result"],
        "settings": {k: defaults[k] for k in ("confidence_threshold", "confidence_smoothing", "trim_ratio", "weak_real_gate",
                                              "weak_real_window_threshold", "fake_vote_gate", "fake_vote_min_windows",
                                              "mouth_motion_check_enabled", "mouth_motion_fake_penalty")},
    }
    return rec


def main():
    out = {}
    for i, (name, confs) in enumerate(scenarios().items()):
        out[f"{name}/plain"] = run(confs, 0.02, 100 + i)
        out[f"{name}/mouth_likely_fake"] = run(confs, 0.02, 100 + i, mouth="likely_fake")
        out[f"{name}/mouth_uncertain"] = run(confs, 0.02, 100 + i, mouth="uncertain")
        out[f"{name}/mouth_check_disabled"] = run(confs, 0.02, 100 + i, mouth="uncertain", mouth_motion_check_enabled=False)
        out[f"{name}/gate010"] = run(confs, 0.02, 100 + i, fake_vote_gate=0.10)          # Settings value (config.py:75)
        out[f"{name}/trimmed"] = run(confs, 0.02, 100 + i, confidence_smoothing="trimmed_mean")
    with open(os.path.join(HERE, "verdict_golden.json"), "w") as fh:
        json.dump(out, fh, indent=0)
    from collections import Counter
    print(len(out), "scenarios;", Counter(v["verdict"] for v in out.values()), Counter(str(v["override_reason"]) for v in out.values()),
          Counter(v["inputs"]["mouth_check_result"] for v in out.values()))


if __name__ == "__main__":
    main()
