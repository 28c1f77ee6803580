"""Golden speaking-alignment scores and mouth-motion checks from the REAL reference (imported from /root/reference, this
container only):

    python tests/golden/make_speech_golden.py

`Predictor._speaking_alignment_score` (app/inference/predictor.py:333-370) and `Predictor._mouth_motion_energy_check`
(:374-419) are called, window by window, exactly as `_predict_long_video` does (:793-800, :1118): float32 `(3,32,96,96)`
crops `/255` + the `_align_audio_chunk` slice of the clip log-mel.  The synthetic uint8 tracks are rebuilt by `make_case`
in the tests, so the fixture stores only the reference's answers.  Output (committed): tests/golden/speech_golden.json
"""
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (seed, n_windows, motion scale, audio mode)
    "random_motion_random_audio": (11, 12, 6.0, "rand"),
    "static_face_loud_audio": (12, 6, 0.0, "loud"),
    "static_face_silent_audio": (13, 6, 0.0, "silent"),
    "speech_like_correlated": (14, 10, 3.0, "corr"),
    "slow_drift_anticorrelated": (15, 8, 1.0, "anti"),
    "short_audio_clamped_tail": (16, 9, 4.0, "short"),
}
T, STRIDE, H, W, F, TA = 32, 8, 96, 96, 80, 128


def make_case(name):
    """-> (track uint8 (n_frames,H,W,3), starts, mel_full float32 (1,80,Ta_full), total_v_frames)"""
    seed, n, scale, mode = CASES[name]
    rng = np.random.RandomState(seed)
    n_frames = T + STRIDE * (n - 1)
    base = rng.randint(40, 216, size=(1, H, W, 3)).astype(np.float64)
    act = np.abs(np.sin(np.arange(n_frames) * 0.35)) + 0.2 * rng.rand(n_frames)      # per-frame "mouth activity"
    noise = rng.randn(n_frames, H, W, 3)
    noise[:, : H // 3] *= 0.2                                                          # less motion in the upper face
    track = np.clip(base + scale * act[:, None, None, None] * noise, 0, 255).astype(np.uint8)
    ta_full = int(n_frames / 15.0 * 100.0) if mode != "short" else int(n_frames / 15.0 * 100.0) - 150
    t_a = np.arange(ta_full)
    act_a = np.interp(t_a / max(1, ta_full - 1), np.arange(n_frames) / max(1, n_frames - 1), act)
    if mode == "rand" or mode == "short":
        mel = -80.0 * rng.rand(F, ta_full)
    elif mode == "loud":
        mel = -15.0 + 3.0 * rng.randn(F, ta_full)
    elif mode == "silent":
        mel = -70.0 + 2.0 * rng.randn(F, ta_full)
    elif mode == "corr":
        mel = -60.0 + 35.0 * act_a[None, :] + 2.0 * rng.randn(F, ta_full)
    else:
        mel = -25.0 - 35.0 * act_a[None, :] + 2.0 * rng.randn(F, ta_full)
    mel = np.clip(mel, -80.0, 0.0).astype(np.float32)[None]
    starts = [STRIDE * i for i in range(n)]
    return track, starts, mel, n_frames


def main():
    sys.path.insert(0, "/root/reference")
    sys.modules.setdefault("librosa", types.ModuleType("librosa"))
    import app.inference.predictor as P

    pred = P.Predictor.__new__(P.Predictor)
    pred.mouth_motion_low_threshold = 0.015       # reference defaults (predictor.py:61-64)
    pred.audio_energy_high_threshold = -25.0
    pred.audio_energy_low_threshold = -50.0
    out = {}
    for name in CASES:
        track, starts, mel, n_frames = make_case(name)
        scores, checks = [], []
        chunks = []
        for s in starts:
            visual = np.ascontiguousarray(track[s:s + T].transpose(3, 0, 1, 2)).astype(np.float32) / 255.0   # video.py:552-556
            audio = pred._align_audio_chunk(mel, s, n_frames)
            chunks.append(visual)
            scores.append(float(P.Predictor._speaking_alignment_score(visual, audio)))
            c = pred._mouth_motion_energy_check(visual, audio)
            checks.append({"audio_energy": float(c["audio_energy"]), "mouth_motion_energy": float(c["mouth_motion_energy"]),
                           "check_result": c["check_result"]})
        agg = pred._aggregate_mouth_motion_check(chunks, starts, mel, n_frames)
        out[name] = {"speaking": scores, "mouth": checks,
                     "aggregate": {k: (v if not isinstance(v, (np.floating, float)) else float(v)) for k, v in agg.items()}}
        print(name, [round(x, 4) for x in scores[:4]], checks[0], agg["check_result"])
    with open(os.path.join(HERE, "speech_golden.json"), "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
