"""Fixture for the log-mel path from an implementation independent of the oracle: torchaudio's MelSpectrogram
configured like librosa>=0.10 + the power_to_db formula.  Run here once:
    python tests/golden/make_logmel_golden.py
Writes tests/golden/logmel_golden.npz (a few strided samples per case, not the full arrays)."""
import os
import sys

import numpy as np
import torch
import torchaudio

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.normpath(os.path.join(HERE, "..", "..")))

CASES = {"noise_20480": (3, 20480, 0.1), "short_1000": (4, 1000, 0.5), "long_50000": (5, 50000, 0.05), "tone_16000": (6, 16000, 0.0)}


def make_pcm(name):
    seed, n, amp = CASES[name]
    g = torch.Generator().manual_seed(seed)
    if name.startswith("tone"):
        t = torch.arange(n, dtype=torch.float32) / 16000.0
        return (0.3 * torch.sin(2 * np.pi * 440.0 * t) + 0.1 * torch.sin(2 * np.pi * 3000.0 * t) + 0.01 * torch.randn(n, generator=g)).numpy()
    return (amp * torch.randn(n, generator=g)).numpy()


def main():
    ms = torchaudio.transforms.MelSpectrogram(sample_rate=16000, n_fft=400, win_length=400, hop_length=160, f_min=0.0, f_max=8000.0,
                                              n_mels=80, power=2.0, center=True, pad_mode="constant", norm="slaney", mel_scale="slaney")
    out = {}
    for name in CASES:
        y = torch.from_numpy(make_pcm(name))
        S = ms(y).numpy()
        db = 10.0 * np.log10(np.maximum(1e-10, S)) - 10.0 * np.log10(np.maximum(1e-10, S.max()))
        db = np.maximum(db, db.max() - 80.0).astype(np.float32)
        out[name + "/shape"] = np.asarray(db.shape)
        flat = db.reshape(-1)
        idx = np.linspace(0, flat.size - 1, 512).astype(np.int64)
        out[name + "/idx"] = idx
        out[name + "/val"] = flat[idx]
        out[name + "/mean"] = np.asarray([flat.mean(), flat.min(), flat.max()])
    np.savez_compressed(os.path.join(HERE, "logmel_golden.npz"), **out)
    print("wrote logmel_golden.npz")


if __name__ == "__main__":
    main()
