"""Golden voice-activity masks from the REAL reference `detect_voice_activity` (app/preprocessing/audio.py:105-245), imported
from /root/reference (this container only) with `librosa.load` stubbed to return the synthetic PCM of `make_pcm`:

    python tests/golden/make_vad_golden.py

Output (committed): tests/golden/vad_golden.json (masks as 0/1 strings + durations).
"""
import json
import os
import sys
import types
from pathlib import Path

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = {"speech_bursts": 21, "continuous_noise": 22, "near_silence": 23, "short_clip": 24, "loud_then_quiet": 25}


def make_pcm(name):
    rng = np.random.RandomState(CASES[name])
    sr = 16000
    if name == "speech_bursts":
        y = 0.002 * rng.randn(5 * sr)
        for a, b in [(0.5, 1.4), (2.0, 2.6), (3.1, 4.4)]:
            seg = slice(int(a * sr), int(b * sr))
            t = np.arange(seg.stop - seg.start) / sr
            y[seg] += 0.2 * np.sin(2 * np.pi * 180 * t) * (0.6 + 0.4 * np.sin(2 * np.pi * 4 * t)) + 0.05 * rng.randn(t.size)
    elif name == "continuous_noise":
        y = 0.1 * rng.randn(3 * sr)
    elif name == "near_silence":
        y = 1e-5 * rng.randn(2 * sr)
    elif name == "short_clip":
        y = 0.05 * rng.randn(1000)
    else:
        y = np.concatenate([0.3 * rng.randn(sr), 1e-3 * rng.randn(2 * sr), 0.3 * rng.randn(sr // 2)])
    return y.astype(np.float32)


def main():
    sys.path.insert(0, "/root/reference")
    lib = types.ModuleType("librosa")
    cur = {}
    lib.load = lambda path, sr=16000: (cur["y"], sr)
    sys.modules["librosa"] = lib
    import app.preprocessing.audio as A
    out = {}
    for name in CASES:
        cur["y"] = make_pcm(name)
        mask, dur = A.detect_voice_activity(Path(f"/nonexistent/{name}.wav"))
        out[name] = {"mask": "".join("1" if m else "0" for m in mask), "duration_sec": float(dur)}
        print(name, len(mask), int(np.sum(mask)), dur)
    with open(os.path.join(HERE, "vad_golden.json"), "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
