"""Golden multi-track decisions from the REAL reference `_predict_long_video` (imported from /root/reference, this container
only):   python tests/golden/make_long_video_golden.py [cuda_confs.json]

Synthetic multi-track videos (seeded uint8 mouth crops -> the reference's fp32 chunks = u8/255, video.py:552-556) go through the
real `Predictor._predict_long_video` with the three preprocessing names monkeypatched and `_infer_confidence`
scripted per (track, window).  Stored per scenario: the generation recipe (seeds, spans, stabilities), the scripted confidences,
the speaking / mouth-motion values the reference computed on the host (so the host logic can be replayed on CPU), and everything
the reference decided (verdict, selected track, selection / confidence-margin flags, per-window winners, timeline).
With `cuda_confs.json` (per-window confidences of this repo's CUDA scorer on the same seeded tracks, dumped by
`scripts/dump_long_video_confs.py` on a B200) the scenario `cuda_scored/*` is generated from those instead of scripted numbers.
Output (committed): tests/golden/long_video_golden.json
"""
import asyncio
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.normpath(os.path.join(HERE, "..", "..")))
sys.modules.setdefault("librosa", types.ModuleType("librosa"))

import torch  # noqa: E402
import app.inference.predictor as P  # noqa: E402
from tests.golden.long_video_synth import make_tracks, make_audio, SCENARIOS  # noqa: E402

KEYS = ["verdict", "is_real", "is_fake", "confidence", "selected_track_id", "selection_uncertain", "selection_margin",
        "turn_taking_detected", "speaker_case", "speaking_tracks_count", "speaking_real_count", "speaking_fake_count", "verdicts",
        "track_policy_verdicts", "conservative_override_applied", "total_chunks_analyzed", "chunks_per_track_max",
        "window_weighted_confidence", "window_fake_vote_ratio", "window_consensus_uncertain", "strict_fake_evidence",
        "confidence_margin_uncertain", "confidence_gap", "sparse_real_guard_applied", "mouth_motion_override_applied",
        "override_reason", "temporal_confidence_drop", "temporal_drift", "first_half_avg_confidence", "second_half_avg_confidence",
        "speaker_timeline"]


def run(spec, confs_by_track):
    tracks_u8, n_frames = make_tracks(spec)
    mel, vad = make_audio(spec, n_frames)
    ref_tracks = []
    for t in tracks_u8:
        chunks = [np.ascontiguousarray(np.transpose(t["crops_u8"][s - t["track_start_frame"]: s - t["track_start_frame"] + 32].astype(np.float32) / 255.0,
                                                    (3, 0, 1, 2))) for s in t["chunk_starts"]]
        ref_tracks.append({k: v for k, v in t.items() if k != "crops_u8"} | {"chunks": chunks})
    P.preprocess_video_tracks_chunked = lambda *a, **k: (ref_tracks, float(spec["fps"]), n_frames)
    P.preprocess_audio = lambda *a, **k: mel
    P.detect_voice_activity = lambda *a, **k: (vad, len(vad) / 100.0)
    p = P.Predictor.__new__(P.Predictor)
    defaults = dict(confidence_threshold=0.5, uncertainty_margin=0.05, confidence_smoothing="median", trim_ratio=0.1,
                    max_tracks=6, chunk_size=32, chunk_stride=8, max_total_frames=None, confidence_margin=0.10,
                    mouth_motion_check_enabled=True, mouth_motion_low_threshold=0.015, mouth_motion_fake_penalty=0.10,
                    audio_energy_high_threshold=-25.0, audio_energy_low_threshold=-50.0, weak_real_gate=0.08,
                    weak_real_window_threshold=0.30, fake_vote_gate=0.15, fake_vote_min_windows=5,
                    refine_margin=0.08, refine_top_k=2, long_video_threshold_sec=3.0, device=torch.device("cpu"))
    for k, v in defaults.items():
        setattr(p, k, v)
    flat = [c for t in ref_tracks for c in confs_by_track[str(t["track_id"])]]
    it = iter(flat)
    p._infer_confidence = lambda v, a: next(it)
    # record what the reference's host statistics returned, in call order
    speak_log, mouth_log = [], []
    orig_speak = P.Predictor._speaking_alignment_score
    orig_mouth = p._aggregate_mouth_motion_check

    def speak(v, a):
        s = float(orig_speak(v, a))
        speak_log.append(s)
        return s

    def mouth(*a, **k):
        r = orig_mouth(*a, **k)
        mouth_log.append(r)
        return r

    p._speaking_alignment_score = speak
    p._aggregate_mouth_motion_check = mouth
    res = asyncio.run(p._predict_long_video(None, None, 0.0))
    rec = {k: res[k] for k in KEYS}
    rec["window_results"] = res["window_results"]
    rec["tracks"] = [{k: t[k] for k in ("track_id", "confidence", "selection_score", "speaking_activity", "is_real", "stability")} for t in res["tracks"]]
    rec["mouth_motion_check"] = res["mouth_motion_check"]
    return rec


def main():
    cuda_confs = None
    if len(sys.argv) > 1:
        with open(sys.argv[1]) as fh:
            cuda_confs = json.load(fh)
    out = {}
    for name, spec in SCENARIOS.items():
        confs = spec["confs"]
        out[f"scripted/{name}"] = {"spec": name, "confs": confs, "expect": run(spec, confs)}
        if cuda_confs and name in cuda_confs:
            out[f"cuda_scored/{name}"] = {"spec": name, "confs": cuda_confs[name], "expect": run(spec, cuda_confs[name])}
    path = os.path.join(HERE, "long_video_golden.json")
    if cuda_confs is None and os.path.exists(path):        # keep previously generated cuda_scored/* entries
        with open(path) as fh:
            for k, v in json.load(fh).items():
                if k.startswith("cuda_scored/"):
                    out.setdefault(k, v)
    with open(path, "w") as fh:
        json.dump(out, fh, indent=0)
    for k, v in out.items():
        e = v["expect"]
        print(k, e["verdict"], "sel", e["selected_track_id"], "sel_unc", e["selection_uncertain"], "margin_unc", e["confidence_margin_uncertain"],
              "turn", e["turn_taking_detected"], e["speaker_case"], e["override_reason"])


if __name__ == "__main__":
    main()
